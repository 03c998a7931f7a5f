// rows.cu -- see rows.cuh.  Row build (rebuild step) and row walk (list steps).
#include "rows.cuh"

#include <algorithm>

#include "kernels.cuh"

namespace b200cv {

namespace {

constexpr int WARPS = ROW_THREADS / 32;

// One pair of a row, branch-free: d = x_partner - x_own.  `valid` lanes that are not next to a
// decision threshold get the sweep's fast form (F = 0 exactly beyond tolerance_l2_max); every
// other lane evaluates l2 = 2^100 and contributes nothing.
template <int EXP, bool ANISO>
__device__ __forceinline__ void row_pair(PairParams const &P, double dx, double dy, double dz,
                                         bool valid, double &F, double &gq, bool &near1,
                                         bool &nearthr)
{
  double ex = dx, ey = dy, ez = dz;
  double const u = pair_l2<ANISO, BC_NONE>(P, ex, ey, ez);
  int const uh = __double2hiint(u);
  near1 = valid && ((unsigned)(uh - 0x3FEE0000) < 0x30000u);
  nearthr = valid && ((unsigned)(uh - P.thr_lo_hi) <= P.thr_width);
  int const keep = (valid && !near1 && !nearthr) ? -1 : 0;
  double const ue =
      __hiloint2double((uh & keep) | (0x46300000 & ~keep), __double2loint(u) & keep);
  switching_fast<EXP, true>(ue, false, P, F, gq);
}

__device__ __forceinline__ void queue_rare(unsigned long long *rare, unsigned rare_cap,
                                           unsigned *rare_count, int atom1, int atom2)
{
  unsigned const slot = atomicAdd(rare_count, 1u);
  if (slot < rare_cap) {
    rare[slot] = ((unsigned long long)(unsigned)atom1 << 32) | (unsigned long long)(unsigned)atom2;
  }
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
  __shared__ int stage[WARPS][64];
  __shared__ int rbuf[WARPS][ROW_BUF];
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
  TileGrid const g = *A.grid;
  PairParams const &P = A.P;
  double const l2_pass = P.l2_max * (1.0 + 1.0e-9);
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
    int const iy_lo = tile_cell(yo - A.cut[1], g.origin[1], g.inv_cell[1], g.dim[1]);
    int const iy_hi = tile_cell(yo + A.cut[1], g.origin[1], g.inv_cell[1], g.dim[1]);
    int const iz_lo = tile_cell(zo - A.cut[2], g.origin[2], g.inv_cell[2], g.dim[2]);
    int const iz_hi = tile_cell(zo + A.cut[2], g.origin[2], g.inv_cell[2], g.dim[2]);
    int const nry = iy_hi - iy_lo + 1;
    int const nspan = nry * (iz_hi - iz_lo + 1);
    double gx = 0.0, gy = 0.0, gz = 0.0;
    int nst = 0;        // staged partners (warp-uniform)
    int rlen = 0;       // buffered row entries (warp-uniform)
    bool first = true;  // the row's own chunk slot is still free

    // the row's entries go out as one chunk -- slot chunk_base + p -- and, for the rare row with
    // more than ROW_BUF entries, further chunks behind the per-row slots
    auto flush = [&](bool last) {
      if (rlen > 0 || (last && first)) {
        unsigned start = 0u, c = A.chunk_base + (unsigned)p;
        if (lane == 0) {
          if (rlen > 0) start = atomicAdd(A.counts + 3, (unsigned)rlen);
          if (!first) c = A.chunk_extra + atomicAdd(A.counts, 1u);
        }
        start = __shfl_sync(0xffffffffu, start, 0);
        c = __shfl_sync(0xffffffffu, c, 0);
        bool const fits = (unsigned long long)start + (unsigned)rlen <= A.cap_idx;
        if (fits) {
          for (int k = lane; k < rlen; k += 32) A.idx[start + k] = R[k];
        }
        if (lane == 0 && c < A.cap_chunks) {
          A.chunks[c] = make_int4(p, (int)start, fits ? rlen : 0, A.dir);
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
      bool near1, nearthr;
      row_pair<EXP, ANISO>(P, dx, dy, dz, valid, F, gq, near1, nearthr);
      bool const canonical = A.self ? (q > p) : (A.canonical != 0);
      fsum += canonical ? F : 0.0;
      gx = fma(gq, dx, gx);
      gy = fma(gq, dy, gy);
      gz = fma(gq, dz, gz);
      if (__builtin_expect((near1 || nearthr) && canonical, 0)) {
        int const nbr_atom = A.nbr.sorted[q];
        if (A.own_is_group1) queue_rare(A.rare, A.rare_cap, A.rare_count, own_atom, nbr_atom);
        else queue_rare(A.rare, A.rare_cap, A.rare_count, nbr_atom, own_atom);
      }
      // in the row: listed by the sign of the shifted function; pairs next to a threshold are
      // listed (or not) by exact_kernel and live in the extras
      bool const listed = valid && !near1 && !nearthr && __double_as_longlong(F) != 0ll;
      unsigned const ml = __ballot_sync(0xffffffffu, listed);
      if (listed) R[rlen + __popc(ml & lt)] = q;
      rlen += __popc(ml);
      npairs += (listed && canonical) ? 1ull : 0ull;
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
          int const ix_lo = tile_cell(xo - xr, g.origin[0], g.inv_cell[0], g.dim[0]);
          int const ix_hi = tile_cell(xo + xr, g.origin[0], g.inv_cell[0], g.dim[0]);
          int const row = (iz * g.dim[1] + iy) * g.dim[0];
          begin = A.nbr.cell_start[row + ix_lo];
          end = A.nbr.cell_start[row + ix_hi + 1];
        }
      }
      int const nsp = min(32, nspan - base);
      for (int s = 0; s < nsp; s++) {
        int const b = __shfl_sync(0xffffffffu, begin, s);
        int const e = __shfl_sync(0xffffffffu, end, s);
        for (int q0 = b; q0 < e; q0 += 32) {
          int const q = q0 + lane;
          bool pass = false;
          if (q < e) {
            double dx = A.nbr.xs[q] - xo, dy = A.nbr.ys[q] - yo, dz = A.nbr.zs[q] - zo;
            double const u = pair_l2<ANISO, BC_NONE>(P, dx, dy, dz);
            pass = (u <= l2_pass) && !(A.self && q == p);
          }
          unsigned const m = __ballot_sync(0xffffffffu, pass);
          if (m == 0u) continue;
          if (pass) S[nst + __popc(m & lt)] = q;
          nst += __popc(m);
          __syncwarp();
          if (nst >= 32) {
            evaluate(32);
            int const rest = nst - 32;
            int moved = 0;
            if (lane < rest) moved = S[32 + lane];
            __syncwarp();
            if (lane < rest) S[lane] = moved;
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
template <int EXP, bool ANISO>
__global__ void __launch_bounds__(ROW_THREADS) row_walk_kernel(const __grid_constant__ RowWalkArgs A)
{
  __shared__ double red[WARPS];
  if (A.gate != nullptr && *A.gate != 0) {
    if (threadIdx.x == 0) A.partials[blockIdx.x] = 0.0;
    return;
  }
  int const lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  PairParams const &P = A.P;
  // work items: the rows of this rank (group1's, then group2's), then the overflow chunks; a
  // warp takes a block of consecutive items (see row_build_kernel)
  unsigned nextra = A.counts[6];
  if (A.chunk_extra + nextra > A.cap_chunks) {
    nextra = A.cap_chunks > A.chunk_extra ? A.cap_chunks - A.chunk_extra : 0u;
  }
  unsigned const n0 = (unsigned)(A.row_end[0] - A.row_begin[0]);
  unsigned const n1 = A.self ? 0u : (unsigned)(A.row_end[1] - A.row_begin[1]);
  unsigned const total = n0 + n1 + nextra;
  double fsum = 0.0;
  unsigned const gwarp = blockIdx.x * WARPS + warp, nwarps = gridDim.x * WARPS;
  unsigned const per = (total + nwarps - 1) / nwarps;
  unsigned const t_first = gwarp * per;
  unsigned const t_last = min(t_first + per, total);
  constexpr int U = 4;  // steps whose loads are issued together
  for (unsigned t = t_first; t < t_last; t++) {
    unsigned const slot = t < n0 ? (unsigned)A.row_begin[0] + t
                                 : (t < n0 + n1 ? A.chunk_base1 + (unsigned)A.row_begin[1] + (t - n0)
                                                : A.chunk_extra + (t - n0 - n1));
    int4 const ch = A.chunks[slot];
    int const p = ch.x, start = ch.y, len = ch.z, dir = ch.w;
    if (len <= 0) continue;
    RowSideView const &O = A.side[A.self ? 0 : dir];
    RowSideView const &N = A.side[A.self ? 0 : 1 - dir];
    double const xo = O.xs[p], yo = O.ys[p], zo = O.zs[p];
    double gx = 0.0, gy = 0.0, gz = 0.0;
    for (int k = 0; k < len; k += 32 * U) {
      int q[U];
      bool valid[U];
      double dx[U], dy[U], dz[U];
#pragma unroll
      for (int u = 0; u < U; u++) {
        int const kk = k + 32 * u + lane;
        valid[u] = kk < len;
        q[u] = A.idx[start + (valid[u] ? kk : 0)];
      }
#pragma unroll
      for (int u = 0; u < U; u++) {
        dx[u] = N.xs[q[u]] - xo;
        dy[u] = N.ys[q[u]] - yo;
        dz[u] = N.zs[q[u]] - zo;
      }
#pragma unroll
      for (int u = 0; u < U; u++) {
        if (k + 32 * u >= len) break;  // warp-uniform
        double F, gq;
        bool near1, nearthr;
        row_pair<EXP, ANISO>(P, dx[u], dy[u], dz[u], valid[u], F, gq, near1, nearthr);
        bool const canonical = A.self ? (q[u] > p) : (dir == 0);
        fsum += canonical ? F : 0.0;
        gx = fma(gq, dx[u], gx);
        gy = fma(gq, dy[u], gy);
        gz = fma(gq, dz[u], gz);
        if (__builtin_expect((near1 || nearthr) && canonical, 0)) {
          int const own_atom = O.sorted[p], nbr_atom = N.sorted[q[u]];
          if (dir == 0) queue_rare(A.rare, A.rare_cap, A.rare_count, own_atom, nbr_atom);
          else queue_rare(A.rare, A.rare_cap, A.rare_count, nbr_atom, own_atom);
        }
      }
    }
    if (A.want_grad) {
      gx = warp_sum(gx);
      gy = warp_sum(gy);
      gz = warp_sum(gz);
      if (lane == 0) {
        int const own_atom = O.sorted[p];
        atomicAdd(&O.grad[own_atom], -P.c_grad[0] * gx);
        atomicAdd(&O.grad[O.gstride + own_atom], -P.c_grad[1] * gy);
        atomicAdd(&O.grad[2 * O.gstride + own_atom], -P.c_grad[2] * gz);
      }
    }
  }
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

int launch_row_walk(const RowWalkArgs &args, int exp_special, bool aniso, cudaStream_t stream)
{
  if (exp_special == 3) {
    if (aniso) row_walk_kernel<3, true><<<ROW_BLOCKS, ROW_THREADS, 0, stream>>>(args);
    else row_walk_kernel<3, false><<<ROW_BLOCKS, ROW_THREADS, 0, stream>>>(args);
  } else {
    if (aniso) row_walk_kernel<0, true><<<ROW_BLOCKS, ROW_THREADS, 0, stream>>>(args);
    else row_walk_kernel<0, false><<<ROW_BLOCKS, ROW_THREADS, 0, stream>>>(args);
  }
  note_launch();
  return (int)cudaGetLastError();
}

}  // namespace b200cv
