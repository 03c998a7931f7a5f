// rows.cu -- see rows.cuh.  Row build (rebuild step) and row walk (list steps).
#include "rows.cuh"

#include <algorithm>

#include "kernels.cuh"

namespace b200cv {

namespace {

constexpr int WARPS = ROW_THREADS / 32;
#ifndef B200CV_ROW_HOIST
#define B200CV_ROW_HOIST 1  // constants of the walk's pair loop pinned in registers
#endif
#ifndef B200CV_ROW_MINB
#define B200CV_ROW_MINB 5  // CTAs per SM the walk's register allocation targets
#endif

// decision thresholds of the pair function as integers (high words of l2), kept in registers
struct RowConsts {
  int thr_lo;       // window around tolerance_l2_max and the root of f0 = tolerance ...
  unsigned thr_w;   // ... [thr_lo, thr_lo + thr_w]
  int l2max_hi;     // high word of tolerance_l2_max
  double s[3];      // isotropic: s[0] = 1/r0^2; per-axis cut-offs: 1/r0 of each axis
  double inv1mt, ntol;
};

template <bool ANISO> __device__ __forceinline__ RowConsts row_consts(PairParams const &P)
{
  RowConsts K;
  K.thr_lo = P.thr_lo_hi;
  K.thr_w = P.thr_width;
  K.l2max_hi = __double2hiint(P.l2_max);
  K.s[0] = ANISO ? P.inv_r0[0] : P.inv_r0sq[0];
  K.s[1] = P.inv_r0[1];
  K.s[2] = P.inv_r0[2];
  K.inv1mt = P.inv1mt;
  K.ntol = P.ntol;
  return K;
}

// One pair of a row, branch-free: d = x_partner - x_own, vm = all ones for a lane that holds a
// pair (all zeros: padding).  A pair next to a decision threshold (`rare`) contributes nothing
// here, nor does one beyond tolerance_l2_max: both evaluate l2 = 2^100 and F is masked to +0
// (gq ~ 2^-400 then: nothing at any scale).  Beyond the threshold window "l2 > tolerance_l2_max"
// can be read off the high word alone, and below it the shifted function is positive
// (pair_is_rare, pair_math.cuh), so no comparison of F is needed.
template <int EXP, bool ANISO>
__device__ __forceinline__ void row_pair(PairParams const &P, RowConsts const &K, double dx,
                                         double dy, double dz, int vm, double &F, double &gq,
                                         bool &rare)
{
  // (pair_l2 of sweep.cuh without boundary conditions, constants from K)
  double u;
  if constexpr (ANISO) {
    double const sx = dx * K.s[0], sy = dy * K.s[1], sz = dz * K.s[2];
    u = fma(sz, sz, fma(sy, sy, sx * sx));
  } else {
    u = fma(dz, dz, fma(dy, dy, dx * dx)) * K.s[0];
  }
  int const uh = __double2hiint(u);
  rare = (((unsigned)(uh - 0x3FEE0000) < 0x30000u) || ((unsigned)(uh - K.thr_lo) <= K.thr_w)) &&
         (vm != 0);
  int const keep = (rare || uh > K.l2max_hi) ? 0 : vm;
  double const ue =
      __hiloint2double((uh & keep) | (0x46300000 & ~keep), __double2loint(u) & keep);
  if constexpr (EXP > 0) {
    double const um1 = ipow<EXP - 1>(ue);
    double const q = fma(um1, ue, 1.0);
    double const r = fast_rcp(q);
    double const Fs = fma(r, K.inv1mt, K.ntol);
    F = __hiloint2double(__double2hiint(Fs) & keep, __double2loint(Fs) & keep);
    gq = um1 * (r * r);
  } else {
    switching_fast<0, true>(ue, false, P, F, gq);
  }
}

// Exact-pair queue of a warp.  Appending to the global queue costs an atomic on ONE counter for
// the whole grid; at a few rare pairs per row that is tens of thousands of serialised
// same-address atomics per step -- more than the pair arithmetic of a list step.  So a warp
// collects its rare pairs in shared memory and appends them in one go (one atomic per RQ_CAP
// entries or per block of rows).
constexpr int RQ_CAP = 64;
struct RareBuf {
  unsigned long long e[RQ_CAP];
};

__device__ __forceinline__ void rare_flush(RareBuf &B, int &n, unsigned long long *rare,
                                           unsigned rare_cap, unsigned *rare_count, int lane)
{
  if (n > 0) {
    __syncwarp();
    unsigned base = 0u;
    if (lane == 0) base = atomicAdd(rare_count, (unsigned)n);
    base = __shfl_sync(0xffffffffu, base, 0);
    for (int k = lane; k < n; k += 32) {
      if (base + (unsigned)k < rare_cap) rare[base + k] = B.e[k];
    }
    __syncwarp();
    n = 0;
  }
}

// warp-wide: lanes with `mine` append (atom1, atom2); n is warp-uniform
__device__ __forceinline__ void rare_push(RareBuf &B, int &n, bool mine, int atom1, int atom2,
                                          unsigned long long *rare, unsigned rare_cap,
                                          unsigned *rare_count, int lane)
{
  unsigned const m = __ballot_sync(0xffffffffu, mine);
  if (m == 0u) return;
  int const add = __popc(m);
  if (n + add > RQ_CAP) rare_flush(B, n, rare, rare_cap, rare_count, lane);
  if (mine) {
    B.e[n + __popc(m & ((1u << lane) - 1u))] =
        ((unsigned long long)(unsigned)atom1 << 32) | (unsigned long long)(unsigned)atom2;
  }
  n += add;
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
  return v;
}

// ---- rebuild: one warp per atom -------------------------------------------------------------------
template <int EXP, bool ANISO>
__global__ void __launch_bounds__(ROW_THREADS) row_build_kernel(const __grid_constant__ RowBuildArgs A)
{
  __shared__ int stage[WARPS][96];  // up to 31 left over + two chunks of 32
  __shared__ int rbuf[WARPS][ROW_BUF];
  __shared__ RareBuf rqs[WARPS];
  __shared__ double red[WARPS];
  __shared__ unsigned long long redn[WARPS];
  if (A.gate != nullptr && *A.gate != 1) {
    if (threadIdx.x == 0) A.partials[blockIdx.x] = 0.0;
    return;
  }
  int const lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned const lt = (1u << lane) - 1u;
  int *const S = stage[warp];
  int *const R = rbuf[warp];
  RareBuf &RQ = rqs[warp];
  int nrq = 0;
  SortGrid const g = *A.grid;
  PairParams const &P = A.P;
  double const l2_pass = P.l2_max * (1.0 + 1.0e-9);
  RowConsts const K = row_consts<ANISO>(P);
  double const cell1 = 1.0 / g.inv_cell[1], cell2 = 1.0 / g.inv_cell[2];
  double fsum = 0.0;
  unsigned long long npairs = 0ull;
  // a warp takes a block of consecutive rows: neighbouring atoms have nearly the same partners,
  // so their coordinates stay in L1 from one row to the next
  int const gwarp = blockIdx.x * WARPS + warp, nwarps = gridDim.x * WARPS;
  int const per = (A.row_end - A.row_begin + nwarps - 1) / nwarps;
  int const p_first = A.row_begin + gwarp * per;
  int const p_last = min(p_first + per, A.row_end);

  for (int p = p_first; p < p_last; p++) {
    double const xo = A.own.xs[p], yo = A.own.ys[p], zo = A.own.zs[p];
    int const own_atom = A.own.sorted[p];
    int const iy_lo = sort_cell(yo - A.cut[1], g.origin[1], g.inv_cell[1], g.dim[1]);
    int const iy_hi = sort_cell(yo + A.cut[1], g.origin[1], g.inv_cell[1], g.dim[1]);
    int const iz_lo = sort_cell(zo - A.cut[2], g.origin[2], g.inv_cell[2], g.dim[2]);
    int const iz_hi = sort_cell(zo + A.cut[2], g.origin[2], g.inv_cell[2], g.dim[2]);
    int const nry = iy_hi - iy_lo + 1;
    int const nspan = nry * (iz_hi - iz_lo + 1);
    // selfCoordNum: a pair belongs to the row of its atom that comes first in cell order, so
    // only partners behind the own position are looked at (the rows of cells in front of the
    // own one are skipped as a whole)
    int const iy_own = sort_cell(yo, g.origin[1], g.inv_cell[1], g.dim[1]);
    int const iz_own = sort_cell(zo, g.origin[2], g.inv_cell[2], g.dim[2]);
    double gx = 0.0, gy = 0.0, gz = 0.0;
    int nst = 0;        // staged partners (warp-uniform)
    int rlen = 0;       // buffered row entries (warp-uniform)
    bool first = true;  // the row's own chunk slot is still free

    // the row's entries go out as one chunk -- slot chunk_base + p -- and, for the rare row with
    // more than ROW_BUF entries, further chunks behind the per-row slots
    auto flush = [&](bool last) {
      if (rlen > 0 || (last && first)) {
        unsigned start = 0u, c = A.chunk_base + (unsigned)p;
        // a whole number of steps of the walk: padded with the far-away point behind the
        // partner coordinates (cellsort.cuh), so chunks start on 128-byte boundaries
        int const plen = (rlen + 31) & ~31;
        if (lane == 0) {
          if (rlen > 0) start = atomicAdd(A.counts + 3, (unsigned)plen);
          if (!first) c = A.chunk_extra + atomicAdd(A.counts, 1u);
        }
        start = __shfl_sync(0xffffffffu, start, 0);
        c = __shfl_sync(0xffffffffu, c, 0);
        bool const fits = (unsigned long long)start + (unsigned)plen <= A.cap_idx;
        if (fits) {
          for (int k = lane; k < plen; k += 32) A.idx[start + k] = k < rlen ? R[k] : A.nbr.n;
        }
        if (lane == 0 && c < A.cap_chunks) {
          A.chunks[c] = make_int4(p, (int)start, fits ? plen : 0, 0);
        }
        __syncwarp();
        rlen = 0;
        first = false;
      }
    };
    // the first nv staged partners at full lane occupancy: value, own gradient, row entries
    auto evaluate = [&](int nv) {
      bool const valid = lane < nv;
      int const q = S[valid ? lane : 0];
      double const dx = A.nbr.xs[q] - xo, dy = A.nbr.ys[q] - yo, dz = A.nbr.zs[q] - zo;
      double F, gq;
      bool rare;
      row_pair<EXP, ANISO>(P, K, dx, dy, dz, valid ? -1 : 0, F, gq, rare);
      fsum += F;
      double const tx = gq * dx, ty = gq * dy, tz = gq * dz;
      gx += tx;
      gy += ty;
      gz += tz;
      if (A.want_grad && valid) {
        // the partner's share (unscaled, in cell order; exact_kernel adds it to the group)
        atomicAdd(&A.nbr.gs[q], tx);
        atomicAdd(&A.nbr.gs[A.nbr.gs_stride + q], ty);
        atomicAdd(&A.nbr.gs[2 * A.nbr.gs_stride + q], tz);
      }
      if (__builtin_expect(__any_sync(0xffffffffu, rare), 0)) {
        // pairs next to a threshold go to exact_kernel
        int const nbr_atom = rare ? A.nbr.sorted[q] : 0;
        rare_push(RQ, nrq, rare, own_atom, nbr_atom, A.rare, A.rare_cap, A.rare_count, lane);
      }
      // in the row: listed by the sign of the shifted function; pairs next to a threshold are
      // listed (or not) by exact_kernel and live in the extras
      bool const listed = valid && !rare && __double_as_longlong(F) != 0ll;
      unsigned const ml = __ballot_sync(0xffffffffu, listed);
      if (listed) R[rlen + __popc(ml & lt)] = q;
      rlen += __popc(ml);
      npairs += listed ? 1ull : 0ull;
      __syncwarp();
      if (rlen > ROW_BUF - 32) flush(false);
    };

    for (int base = 0; base < nspan; base += 32) {
      // one row of cells per lane: the x-range the atom can reach there
      int begin = 0, end = 0;
      int const k = base + lane;
      if (k < nspan) {
        int const iy = iy_lo + k % nry, iz = iz_lo + k / nry;
        // the cross-section of the row of cells (a little wider than the binning makes it; the
        // outermost cells also hold everything beyond the grid)
        double const ylo = iy == 0 ? -1.0e300 : g.origin[1] + ((double)iy - 1.0e-9) * cell1;
        double const yhi =
            iy == g.dim[1] - 1 ? 1.0e300 : g.origin[1] + ((double)(iy + 1) + 1.0e-9) * cell1;
        double const zlo = iz == 0 ? -1.0e300 : g.origin[2] + ((double)iz - 1.0e-9) * cell2;
        double const zhi =
            iz == g.dim[2] - 1 ? 1.0e300 : g.origin[2] + ((double)(iz + 1) + 1.0e-9) * cell2;
        double const sy = fmax(fmax(ylo - yo, yo - yhi), 0.0) * P.inv_r0[1];
        double const sz = fmax(fmax(zlo - zo, zo - zhi), 0.0) * P.inv_r0[2];
        double const r2 = fma(sz, sz, sy * sy);
        if (r2 <= l2_pass) {
          double const xr = sqrt(l2_pass - r2) / P.inv_r0[0] * (1.0 + 1.0e-9);
          int const ix_lo = sort_cell(xo - xr, g.origin[0], g.inv_cell[0], g.dim[0]);
          int const ix_hi = sort_cell(xo + xr, g.origin[0], g.inv_cell[0], g.dim[0]);
          int const row = (iz * g.dim[1] + iy) * g.dim[0];
          begin = A.nbr.cell_start[row + ix_lo];
          end = A.nbr.cell_start[row + ix_hi + 1];
          if (A.self) {
            bool const before = iz < iz_own || (iz == iz_own && iy < iy_own);
            if (before) end = begin;
            else if (iz == iz_own && iy == iy_own) begin = max(begin, p + 1);
          }
        }
      }
      int const nsp = min(32, nspan - base);
      for (int s = 0; s < nsp; s++) {
        int const b = __shfl_sync(0xffffffffu, begin, s);
        int const e = __shfl_sync(0xffffffffu, end, s);
        // two chunks of 32 candidates per trip: their loads are independent
        for (int q0 = b; q0 < e; q0 += 64) {
          int const qa = q0 + lane, qb = q0 + 32 + lane;
          bool pa = false, pb = false;
          double ua = 0.0, ub = 0.0;
          if (qa < e) {
            double dx = A.nbr.xs[qa] - xo, dy = A.nbr.ys[qa] - yo, dz = A.nbr.zs[qa] - zo;
            ua = pair_l2<ANISO, BC_NONE>(P, dx, dy, dz);
            pa = true;
          }
          if (qb < e) {
            double dx = A.nbr.xs[qb] - xo, dy = A.nbr.ys[qb] - yo, dz = A.nbr.zs[qb] - zo;
            ub = pair_l2<ANISO, BC_NONE>(P, dx, dy, dz);
            pb = true;
          }
          pa = pa && (ua <= l2_pass);
          pb = pb && (ub <= l2_pass);
          unsigned const ma = __ballot_sync(0xffffffffu, pa);
          unsigned const mb = __ballot_sync(0xffffffffu, pb);
          if ((ma | mb) == 0u) continue;
          if (pa) S[nst + __popc(ma & lt)] = qa;
          nst += __popc(ma);
          if (pb) S[nst + __popc(mb & lt)] = qb;
          nst += __popc(mb);
          __syncwarp();
          while (nst >= 32) {
            evaluate(32);
            int const rest = nst - 32;  // < 64
            int m0 = 0, m1 = 0;
            if (lane < rest) m0 = S[32 + lane];
            if (lane + 32 < rest) m1 = S[64 + lane];
            __syncwarp();
            if (lane < rest) S[lane] = m0;
            if (lane + 32 < rest) S[32 + lane] = m1;
            nst = rest;
            __syncwarp();
          }
        }
      }
    }
    if (nst > 0) evaluate(nst);
    flush(true);
    if (A.want_grad) {
      gx = warp_sum(gx);
      gy = warp_sum(gy);
      gz = warp_sum(gz);
      if (lane == 0) {
        atomicAdd(&A.own.grad[own_atom], -P.c_grad[0] * gx);
        atomicAdd(&A.own.grad[A.own.gstride + own_atom], -P.c_grad[1] * gy);
        atomicAdd(&A.own.grad[2 * A.own.gstride + own_atom], -P.c_grad[2] * gz);
      }
    }
  }

  rare_flush(RQ, nrq, A.rare, A.rare_cap, A.rare_count, lane);
  fsum = warp_sum(fsum);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) npairs += __shfl_xor_sync(0xffffffffu, npairs, d);
  if (lane == 0) {
    red[warp] = fsum;
    redn[warp] = npairs;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    unsigned long long np = 0ull;
    for (int w = 0; w < WARPS; w++) {
      s += red[w];
      np += redn[w];
    }
    A.partials[blockIdx.x] = s;
    if (np) atomicAdd(reinterpret_cast<unsigned long long *>(A.counts + 4), np);
  }
}

// ---- list step: one warp per row chunk ------------------------------------------------------------
// The walk is a stream of dependent loads -- chunk header -> partner indices -> partner
// coordinates -- over a 50 MB index array that comes from HBM every step (at BASELINE config 3),
// so each warp runs a software pipeline over its block of rows: while row t is evaluated, the
// header of row t + 2 is being loaded and the indices (and own coordinates) of row t + 1 are on
// their way into a shared-memory buffer (cp.async, 16 bytes per lane: chunks start on 128-byte
// boundaries), two buffers per warp; inside a row the partner coordinates of the next two steps
// are loaded while the current two are evaluated.  A chunk is a whole number of steps (the build
// pads it with the index of the far-away point behind the sorted coordinates), so the pair loop
// has no lane masks, and it is straight-line code: one vote per two steps sends the few pairs
// that have drifted next to a decision threshold to the queue.
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src)
{
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)),
               "l"(gmem_src)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait()
{
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// sums of three per-lane values over the warp in 12 shuffles instead of 30: the halves of the warp
// swap what the other half keeps.  Lane 0 returns the sum of a, lane 8 of b, lane 16 of c.
__device__ __forceinline__ double warp_sum3(double a, double b, double c, int lane)
{
  bool const h16 = (lane & 16) != 0, h8 = (lane & 8) != 0;
  double k0 = h16 ? c : a, k1 = h16 ? 0.0 : b;
  double const s0 = h16 ? a : c, s1 = h16 ? b : 0.0;
  k0 += __shfl_xor_sync(0xffffffffu, s0, 16);
  k1 += __shfl_xor_sync(0xffffffffu, s1, 16);
  double k = h8 ? k1 : k0;
  double const s = h8 ? k0 : k1;
  k += __shfl_xor_sync(0xffffffffu, s, 8);
  k += __shfl_xor_sync(0xffffffffu, k, 4);
  k += __shfl_xor_sync(0xffffffffu, k, 2);
  k += __shfl_xor_sync(0xffffffffu, k, 1);
  return k;
}

template <int EXP, bool ANISO>
__global__ void __launch_bounds__(ROW_THREADS, B200CV_ROW_MINB) row_walk_kernel(const __grid_constant__ RowWalkArgs A)
{
  __shared__ double red[WARPS];
  __shared__ __align__(16) int sidx[WARPS][2][ROW_BUF];
  __shared__ RareBuf rqs[WARPS];
  // (the grid is one wave of CTAs; the final sum reads ROW_BLOCKS partial sums)
  if (blockIdx.x == 0) {
    for (unsigned b = gridDim.x + threadIdx.x; b < (unsigned)ROW_BLOCKS; b += ROW_THREADS) A.partials[b] = 0.0;
  }
  if (A.gate != nullptr && *A.gate != 0) {
    if (threadIdx.x == 0) A.partials[blockIdx.x] = 0.0;
    return;
  }
  int const lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  PairParams const &P = A.P;
  // work items: the rows of this rank, then the overflow chunks; a warp takes a block of
  // consecutive items (see row_build_kernel)
  unsigned nextra = A.counts[6];
  if (A.chunk_extra + nextra > A.cap_chunks) {
    nextra = A.cap_chunks > A.chunk_extra ? A.cap_chunks - A.chunk_extra : 0u;
  }
  unsigned const n0 = (unsigned)(A.row_end - A.row_begin);
  unsigned const total = n0 + nextra;
  RowSideView const &O = A.own;
  RowSideView const &N = A.nbr;
  double fsum = 0.0;
  unsigned const gwarp = blockIdx.x * WARPS + warp, nwarps = gridDim.x * WARPS;
  unsigned const per = (total + nwarps - 1) / nwarps;
  unsigned const t_first = gwarp * per;
  unsigned const t_last = min(t_first + per, total);
  // the constants of the pair loop, through shared memory: values the assembler cannot trace
  // back to the parameter bank stay in registers instead of being re-read on every trip
  __shared__ RowConsts kshared;
  if (B200CV_ROW_HOIST) {
    if (threadIdx.x == 0) kshared = row_consts<ANISO>(P);
    __syncthreads();
  }
  RowConsts const K = B200CV_ROW_HOIST ? kshared : row_consts<ANISO>(P);
  RareBuf &RQ = rqs[warp];
  int nrq = 0;

  auto header = [&](unsigned t) -> int4 {
    if (t >= t_last) return make_int4(0, 0, 0, 0);
    unsigned const slot = t < n0 ? (unsigned)A.row_begin + t : A.chunk_extra + (t - n0);
    return A.chunks[slot];
  };
  // indices of a chunk into one of the warp's two buffers (asynchronously)
  auto prefetch = [&](int4 const &ch, int buf) {
    int const len = min(ch.z, ROW_BUF);
    for (int k = 4 * lane; k < len; k += 128) cp_async16(&sidx[warp][buf][k], A.idx + ch.y + k);
    cp_async_commit();
  };

  int4 h1 = header(t_first), h2 = header(t_first + 1);
  prefetch(h1, 0);
  // own coordinates (and atom) of the coming row, loaded one row ahead like its indices
  double nxo = 0.0, nyo = 0.0, nzo = 0.0;
  int nown = 0;
  if (t_first < t_last) {
    nxo = O.xs[h1.x];
    nyo = O.ys[h1.x];
    nzo = O.zs[h1.x];
    nown = O.sorted[h1.x];
  }
  for (unsigned t = t_first; t < t_last; t++) {
    int4 const ch = h1;
    double const xo = nxo, yo = nyo, zo = nzo;
    int const own_atom = nown;
    int const buf = (int)((t - t_first) & 1u);
    h1 = h2;
    h2 = header(t + 2);
    prefetch(h1, buf ^ 1);  // (an empty group past the last row)
    if (t + 1 < t_last) {
      nxo = O.xs[h1.x];
      nyo = O.ys[h1.x];
      nzo = O.zs[h1.x];
      nown = O.sorted[h1.x];
    }
    cp_async_wait<1>();  // this row's indices have landed (the next row's may still be in flight)
    __syncwarp();
    int const len = min(ch.z, ROW_BUF);  // a multiple of 32
    if (len > 0) {
      const double *const nx = N.xs, *const ny = N.ys, *const nz = N.zs;
      double *const gsx = N.gs, *const gsy = N.gs + N.gs_stride, *const gsz = N.gs + 2 * N.gs_stride;
      int const nbr_n = N.n;
      const int *const qp = sidx[warp][buf] + lane;
      double gx = 0.0, gy = 0.0, gz = 0.0, rsum = 0.0;
      auto load = [&](int k, double &x, double &y, double &z) {
        int const q = qp[k];
        x = nx[q];
        y = ny[q];
        z = nz[q];
      };
      // one pair: the own atom's share stays in registers, the partner's (unscaled, in cell
      // order; exact_kernel adds it to the group) goes out as three reductions
      auto eval = [&](int k, double x, double y, double z, bool &rare) {
        double const dx = x - xo, dy = y - yo, dz = z - zo;
        double F, gq;
        row_pair<EXP, ANISO>(P, K, dx, dy, dz, -1, F, gq, rare);
        rsum += F;
        double const tx = gq * dx, ty = gq * dy, tz = gq * dz;
        gx += tx;
        gy += ty;
        gz += tz;
        int const q = qp[k];
        if (A.want_grad && q != nbr_n) {  // (not the padding: every row would add to one address)
          atomicAdd(&gsx[q], tx);
          atomicAdd(&gsy[q], ty);
          atomicAdd(&gsz[q], tz);
        }
      };
      // a pair next to a threshold goes to exact_kernel
      auto queue = [&](int k, bool rare) {
        int const q = qp[k];
        int const nbr_atom = rare ? N.sorted[q] : 0;
        rare_push(RQ, nrq, rare, own_atom, nbr_atom, A.rare, A.rare_cap, A.rare_count, lane);
      };
      double ax0, ay0, az0, ax1 = 0.0, ay1 = 0.0, az1 = 0.0;
      load(0, ax0, ay0, az0);
      if (len > 32) load(32, ax1, ay1, az1);
      int k = 0;
      for (; k + 64 <= len; k += 64) {
        double bx0, by0, bz0, bx1, by1, bz1;  // (read only where they were loaded)
        if (k + 64 < len) load(k + 64, bx0, by0, bz0);
        if (k + 96 < len) load(k + 96, bx1, by1, bz1);
        bool r0, r1;
        eval(k, ax0, ay0, az0, r0);
        eval(k + 32, ax1, ay1, az1, r1);
        if (__builtin_expect(__any_sync(0xffffffffu, r0 || r1), 0)) {
          queue(k, r0);
          queue(k + 32, r1);
        }
        ax0 = bx0, ay0 = by0, az0 = bz0;
        ax1 = bx1, ay1 = by1, az1 = bz1;
      }
      if (k < len) {
        bool r0;
        eval(k, ax0, ay0, az0, r0);
        if (__builtin_expect(__any_sync(0xffffffffu, r0), 0)) queue(k, r0);
      }
      fsum += rsum;
      if (A.want_grad) {
        double const g = warp_sum3(gx, gy, gz, lane);
        if ((lane & 7) == 0 && lane < 24) {
          int const c = lane >> 3;
          atomicAdd(&O.grad[(size_t)c * O.gstride + own_atom], -P.c_grad[c] * g);
        }
      }
    }
    __syncwarp();  // the buffer is free for the row after next
  }
  cp_async_wait<0>();
  rare_flush(RQ, nrq, A.rare, A.rare_cap, A.rare_count, lane);
  fsum = warp_sum(fsum);
  if (lane == 0) red[warp] = fsum;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < WARPS; w++) s += red[w];
    A.partials[blockIdx.x] = s;
  }
}

}  // namespace

int rows_reserve(RowList &R, unsigned entries, unsigned chunks)
{
  cudaError_t e;
  if (entries > R.cap_idx || !R.d_idx) {
    cudaFree(R.d_idx);
    R.d_idx = nullptr;
    R.cap_idx = 0;
    if ((e = cudaMalloc(&R.d_idx, sizeof(int) * (size_t)std::max(entries, 1u))) != cudaSuccess) {
      return (int)e;
    }
    R.cap_idx = entries;
  }
  if (chunks > R.cap_chunks || !R.d_chunks) {
    cudaFree(R.d_chunks);
    R.d_chunks = nullptr;
    R.cap_chunks = 0;
    if ((e = cudaMalloc(&R.d_chunks, sizeof(int4) * (size_t)std::max(chunks, 1u))) != cudaSuccess) {
      return (int)e;
    }
    R.cap_chunks = chunks;
  }
  return 0;
}

void rows_free(RowList &R)
{
  cudaFree(R.d_idx);
  cudaFree(R.d_chunks);
  R = RowList();
}

int launch_row_build(const RowBuildArgs &args, int exp_special, bool aniso, cudaStream_t stream)
{
  if (exp_special == 3) {
    if (aniso) row_build_kernel<3, true><<<ROW_BLOCKS, ROW_THREADS, 0, stream>>>(args);
    else row_build_kernel<3, false><<<ROW_BLOCKS, ROW_THREADS, 0, stream>>>(args);
  } else {
    if (aniso) row_build_kernel<0, true><<<ROW_BLOCKS, ROW_THREADS, 0, stream>>>(args);
    else row_build_kernel<0, false><<<ROW_BLOCKS, ROW_THREADS, 0, stream>>>(args);
  }
  note_launch();
  return (int)cudaGetLastError();
}

// one wave of CTAs: what the device holds at once (rows are dealt out in equal blocks; a second,
// partly filled wave would run at a fraction of the occupancy)
template <int EXP, bool ANISO> static int walk_grid()
{
  static int grid = 0;
  if (grid == 0) {
    int dev = 0, sms = 148, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, row_walk_kernel<EXP, ANISO>, ROW_THREADS, 0);
    grid = std::max(1, std::min(ROW_BLOCKS, sms * std::max(per_sm, 1)));
  }
  return grid;
}

template <int EXP, bool ANISO> static void walk_launch(const RowWalkArgs &args, cudaStream_t stream)
{
  row_walk_kernel<EXP, ANISO><<<walk_grid<EXP, ANISO>(), ROW_THREADS, 0, stream>>>(args);
}

int launch_row_walk(const RowWalkArgs &args, int exp_special, bool aniso, cudaStream_t stream)
{
  if (exp_special == 3) {
    if (aniso) walk_launch<3, true>(args, stream);
    else walk_launch<3, false>(args, stream);
  } else {
    if (aniso) walk_launch<0, true>(args, stream);
    else walk_launch<0, false>(args, stream);
  }
  note_launch();
  return (int)cudaGetLastError();
}

}  // namespace b200cv
