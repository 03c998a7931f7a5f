// sweep.cuh -- the all-pairs tile kernel of coordNum / selfCoordNum (sm_100a).
//
// Replaces the O(N1*N2) double loop colvar::coordnum::main_loop
// (src/colvarcomp_coordnums.cpp:187-248) and selfcoordnum_sequential_loop (:478-523).
//
// Decomposition
//   work item   = (block of IB = W*32*R group1 atoms) x (chunk of <= JCMAX group2 atoms);
//                 one CTA of W warps per item, items listed in a device table so that
//                 the same kernel serves coordNum, the upper triangle of selfCoordNum and
//                 any multi-GPU shard of rows.
//   group1 side : each lane keeps R atoms (coordinates + 3 gradient accumulators) in
//                 registers for the whole item.
//   group2 side : the chunk's SoA coordinates are staged into shared memory once per CTA
//                 with three TMA bulk copies (cp.async.bulk, mbarrier complete_tx).
//   tile        : 32 group2 atoms.  At step t lane l works on atom (l+t)%32 of the tile,
//                 read conflict-free from shared memory; its 3 group2 gradient
//                 accumulators travel with the atom by one SHFL rotation per step, so
//                 after 32 steps lane l owns the complete warp-level sum for atom l with
//                 no FP64 work spent on reductions.
//   group2 grads: warps walk the chunk's tiles in a staggered order and add their
//                 per-tile sums into block-private shared-memory accumulators; one
//                 RED.ADD.F64 per (CTA, atom, axis) goes to global memory at the end.
//   value       : per-lane sum -> warp shuffle tree -> one partial per CTA (summed in a
//                 fixed order by finalize_kernel: no atomics on the scalar).
//   pair list   : on rebuild sweeps each lane records one 32-bit mask per register atom
//                 and tile; masks are compacted into a global (i,j) list by a warp
//                 prefix sum and one atomicAdd per warp and tile (stream compaction).
//   rare pairs  : pairs within 2^-4 of l2 = 1 (or next to a pair-list threshold, or with an
//                 ambiguous image in a mixed / triclinic cell) are appended to a queue and
//                 evaluated in the reference's operation order by exact_kernel; they
//                 contribute nothing here.
//   cells       : PBC = BC_NONE, BC_ORTHO (minimum image per axis, inside pair_l2) or BC_GENERAL
//                 (mixed / triclinic: general_image, pair_math.cuh -- the reference's search
//                 over the 27 neighbouring images, decided on 13 dot-product scores).
//
// FP64-pipe cost of the specialised n=6/m=12 isotropic pair: 3 (diff) + 3 (d2) + 1 (l2)
// + 2 (l2^2, 1+l2^3) + 3 (reciprocal refinement) + 2 (r^2, gq) + 1 (F sum) + 6 (two
// gradient FMAs per axis) = 21 DFMA-class instructions, vs 39 algorithmic flops.
#pragma once
#include "pair_math.cuh"

namespace b200cv {

#ifndef B200CV_JCMAX
#define B200CV_JCMAX 1024
#endif
constexpr int JCMAX = B200CV_JCMAX;
#ifndef B200CV_GEN_UNROLL
#define B200CV_GEN_UNROLL 2  // image searches (general cells) interleaved per loop trip
#endif
#ifndef B200CV_TUNROLL
#define B200CV_TUNROLL 1  // unroll factor of the 32-step rotation loop of a tile
#endif  // group2 atoms per chunk (24 B coordinates + 24 B accumulators each)

struct SweepItem {
  int i0;    // first group1 atom of the block
  int j0;    // first group2 atom of the chunk (multiple of 32)
  int jn;    // atoms in the chunk (<= JCMAX)
  int diag;  // selfCoordNum only: chunk overlaps the block's rows -> apply j > i
};

struct SweepArgs {
  const double *x1, *y1, *z1;  // group1 planes
  const double *x2, *y2, *z2;  // group2 planes, readable up to the next multiple of 32
  double *g1x, *g1y, *g1z;
  double *g2x, *g2y, *g2z;
  const SweepItem *items;
  double *partials;  // one per CTA
  unsigned long long *rare;
  unsigned int rare_cap;
  unsigned int *rare_count;
  unsigned long long *pl;
  unsigned long long pl_cap;
  unsigned long long *pl_count;
  int i_end;      // group1 atoms with index >= i_end do not exist
  int want_grad;  // 0: value only (no gradient flush)
  // captured (graph) pair-list evaluations enqueue the rebuild sweep AND the list walk; each
  // kernel runs only when the device word *gate equals gate_on (null: always)
  const int *gate;
  int gate_on;
  PairParams P;
};

// ---- TMA bulk copy + mbarrier (PTX) --------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init()
{
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// ---- pair geometry -------------------------------------------------------------------------
// UNIT: plain squared distance (distanceInv), no cutoff scaling
template <bool ANISO, int PBC, bool UNIT = false>
__device__ __forceinline__ double pair_l2(PairParams const &P, double &dx, double &dy, double &dz)
{
  if constexpr (PBC == BC_ORTHO) {
    // position_distance, orthogonal cell (src/colvars_system.h:176-184): the image index is
    // computed in the reference's operation order so that ties break identically
    double const shx = floor(__dadd_rn(__dmul_rn(P.recip[0][0], dx), 0.5));
    double const shy = floor(__dadd_rn(__dmul_rn(P.recip[1][1], dy), 0.5));
    double const shz = floor(__dadd_rn(__dmul_rn(P.recip[2][2], dz), 0.5));
    dx = fma(-shx, P.cell[0][0], dx);
    dy = fma(-shy, P.cell[1][1], dy);
    dz = fma(-shz, P.cell[2][2], dz);
  }
  if constexpr (UNIT) {
    return fma(dz, dz, fma(dy, dy, dx * dx));
  } else if constexpr (ANISO) {
    double const sx_ = dx * P.inv_r0[0];
    double const sy_ = dy * P.inv_r0[1];
    double const sz_ = dz * P.inv_r0[2];
    return fma(sz_, sz_, fma(sy_, sy_, sx_ * sx_));
  } else {
    return fma(dz, dz, fma(dy, dy, dx * dx)) * P.inv_r0sq[0];
  }
}

// ---- one 32-atom tile ---------------------------------------------------------------------
// Two ways of keeping rare pairs out of the fast sums:
//  * SELECT (masked tiles, pair-list sweeps, runtime exponents): per pair, test l2 and switch
//    the pair off before it is evaluated;
//  * UNDO_PL: the same idea for pair-list rebuild sweeps; the decision "listed or not" is the
//    sign of the shifted function, recorded per pair in the tile's bit masks, and the cold
//    block removes exactly the recorded contributions of pairs inside a decision window.
//  * UNDO (specialised exponents, no pair list, interior tiles -- the benchmark path): per pair
//    only fold the high word of l2 into a running minimum (2 integer instructions);
//    evaluate and accumulate every pair; if the step saw a rare pair (about 1 step in 5000 at
//    water density) a cold block re-evaluates it, subtracts it again and queues it.  The
//    specialised form 1/(1+l2^n') is finite and smooth at l2 = 1, so adding and removing a
//    pair costs one rounding of the running sums and nothing else.  The FP64 pipe holds the
//    issue port for 2-3 cycles per instruction and nothing co-issues with it on sm_100
//    (tools/ubench_fp64.cu), so every integer instruction removed from the pair is time.
template <int EXP, bool ANISO, int PBC, bool PL, int R, bool MASKED>
__device__ __forceinline__ void sweep_tile(SweepArgs const &A, SweepItem const &item, int tile,
                                           int lane, int ibase, const double *__restrict__ sx,
                                           const double *__restrict__ sy,
                                           const double *__restrict__ sz, double *ax, double *ay,
                                           double *az, const double (&xi)[R], const double (&yi)[R],
                                           const double (&zi)[R], double (&a1x)[R],
                                           double (&a1y)[R], double (&a1z)[R], double &fsum,
                                           const double *__restrict__ tritab)
{
  // general cells take the select path: an ambiguous image decision is one more reason to
  // queue a pair, known only per pair
  constexpr bool UNDO = (EXP > 0) && !PL && !MASKED && (PBC != BC_GENERAL);
  constexpr bool UNDO_PL = (EXP > 0) && PL && !MASKED && (PBC != BC_GENERAL);
  PairParams const &P = A.P;
  double b2x = 0.0, b2y = 0.0, b2z = 0.0;
  unsigned plm[R];
#pragma unroll
  for (int k = 0; k < R; k++) plm[k] = 0u;
  int const tbase = tile * 32;
  const double *__restrict__ px = sx + tbase;
  const double *__restrict__ py = sy + tbase;
  const double *__restrict__ pz = sz + tbase;

  constexpr int TUNROLL = B200CV_TUNROLL;
#pragma unroll TUNROLL
  for (int t = 0; t < 32; t++) {
    int const jj = (lane + t) & 31;
    double const xj = px[jj];
    double const yj = py[jj];
    double const zj = pz[jj];
    int const jloc = tbase + jj;
    int const jglob = item.j0 + jloc;
    unsigned rmask = 0u;
    if constexpr (UNDO) {
      unsigned nearest = 0xffffffffu;  // min over the step of (hi(l2) - hi(1 - 2^-4)), unsigned
#pragma unroll
      for (int k = 0; k < R; k++) {
        double dx = xj - xi[k];
        double dy = yj - yi[k];
        double dz = zj - zi[k];
        double const u = pair_l2<ANISO, PBC>(P, dx, dy, dz);
        nearest = min(nearest, (unsigned)(__double2hiint(u) - 0x3FEE0000));
        double F, gq;
        switching_fast<EXP, false>(u, false, P, F, gq);
        fsum += F;
        a1x[k] = fma(gq, dx, a1x[k]);
        b2x = fma(gq, dx, b2x);
        a1y[k] = fma(gq, dy, a1y[k]);
        b2y = fma(gq, dy, b2y);
        a1z[k] = fma(gq, dz, a1z[k]);
        b2z = fma(gq, dz, b2z);
      }
      if (__builtin_expect(nearest < 0x30000u, 0)) {
        // cold: some pair of this step has |l2 - 1| < 2^-4.  Take it out again.
#pragma unroll
        for (int k = 0; k < R; k++) {
          double dx = xj - xi[k];
          double dy = yj - yi[k];
          double dz = zj - zi[k];
          double const u = pair_l2<ANISO, PBC>(P, dx, dy, dz);
          if (near_one(u)) {
            double F, gq;
            switching_fast<EXP, false>(u, false, P, F, gq);
            fsum -= F;
            a1x[k] = fma(-gq, dx, a1x[k]);
            b2x = fma(-gq, dx, b2x);
            a1y[k] = fma(-gq, dy, a1y[k]);
            b2y = fma(-gq, dy, b2y);
            a1z[k] = fma(-gq, dz, a1z[k]);
            b2z = fma(-gq, dz, b2z);
            rmask |= 1u << k;
          }
        }
      }
    } else if constexpr (UNDO_PL) {
      // pair-list sweep, interior tile: the yes/no decision is the sign of the shifted
      // function (outside the decision window that is also "l2 <= l2_max"); it is recorded
      // in plm, so the cold block knows exactly what to take out again.
      unsigned near1 = 0xffffffffu, nearthr = 0xffffffffu;
#pragma unroll
      for (int k = 0; k < R; k++) {
        double dx = xj - xi[k];
        double dy = yj - yi[k];
        double dz = zj - zi[k];
        double const u = pair_l2<ANISO, PBC>(P, dx, dy, dz);
        int const uh = __double2hiint(u);
        near1 = min(near1, (unsigned)(uh - 0x3FEE0000));
        nearthr = min(nearthr, (unsigned)(uh - P.thr_lo_hi));
        double Fh, gqh;
        switching_shifted<EXP>(u, P, Fh, gqh);
        bool const keep = __double_as_longlong(Fh) > 0ll;
        double const F = keep ? Fh : 0.0;
        double const gq = keep ? gqh : 0.0;
        plm[k] |= (keep ? 1u : 0u) << t;
        fsum += F;
        a1x[k] = fma(gq, dx, a1x[k]);
        b2x = fma(gq, dx, b2x);
        a1y[k] = fma(gq, dy, a1y[k]);
        b2y = fma(gq, dy, b2y);
        a1z[k] = fma(gq, dz, a1z[k]);
        b2z = fma(gq, dz, b2z);
      }
      if (__builtin_expect((near1 < 0x30000u) || (nearthr <= P.thr_width), 0)) {
#pragma unroll
        for (int k = 0; k < R; k++) {
          double dx = xj - xi[k];
          double dy = yj - yi[k];
          double dz = zj - zi[k];
          double const u = pair_l2<ANISO, PBC>(P, dx, dy, dz);
          if (pair_is_rare<true>(u, P)) {
            if ((plm[k] >> t) & 1u) {
              double Fh, gqh;
              switching_shifted<EXP>(u, P, Fh, gqh);
              fsum -= Fh;
              a1x[k] = fma(-gqh, dx, a1x[k]);
              b2x = fma(-gqh, dx, b2x);
              a1y[k] = fma(-gqh, dy, a1y[k]);
              b2y = fma(-gqh, dy, b2y);
              a1z[k] = fma(-gqh, dz, a1z[k]);
              b2z = fma(-gqh, dz, b2z);
              plm[k] &= ~(1u << t);
            }
            rmask |= 1u << k;
          }
        }
      }
    } else {
      // general cells: the image search is ~300 instructions per pair; unrolled R times (and
      // again for the masked variant) the loop body outgrows the instruction cache (ncu: the
      // warps then wait for instructions 2 cycles out of 3).  So the R searches run as a
      // rolled loop over a small per-thread array (local memory, L1-resident), and only the
      // light part of the pair stays unrolled.
      double gdx[PBC == BC_GENERAL ? R : 1], gdy[PBC == BC_GENERAL ? R : 1],
          gdz[PBC == BC_GENERAL ? R : 1];
      unsigned ambmask = 0u;
      if constexpr (PBC == BC_GENERAL) {
#pragma unroll
        for (int k = 0; k < R; k++) {
          gdx[k] = xj - xi[k];
          gdy[k] = yj - yi[k];
          gdz[k] = zj - zi[k];
        }
        constexpr int GEN_UNROLL = B200CV_GEN_UNROLL;
#pragma unroll GEN_UNROLL
        for (int k = 0; k < R; k++) {
          double ex = gdx[k], ey = gdy[k], ez = gdz[k];
          bool const amb = general_image(P, tritab, ex, ey, ez);
          gdx[k] = ex;
          gdy[k] = ey;
          gdz[k] = ez;
          ambmask |= (amb ? 1u : 0u) << k;
        }
      }
#pragma unroll
      for (int k = 0; k < R; k++) {
        double dx, dy, dz;
        bool ambiguous = false;
        if constexpr (PBC == BC_GENERAL) {
          dx = gdx[k];
          dy = gdy[k];
          dz = gdz[k];
          ambiguous = (ambmask >> k) & 1u;
        } else {
          dx = xj - xi[k];
          dy = yj - yi[k];
          dz = zj - zi[k];
        }
        double const u =
            pair_l2<ANISO, (PBC == BC_GENERAL ? BC_NONE : PBC), (EXP < 0)>(P, dx, dy, dz);
        // branch-free: a rare pair contributes nothing here and is remembered in a mask, so
        // that the R pairs of this step stay one basic block for the instruction scheduler
        bool rare = ((EXP < 0) ? false : pair_is_rare<PL>(u, P)) || ambiguous;
        bool off = rare;
        if constexpr (MASKED) {
          int const iglob = ibase + k * 32;
          bool const valid =
              (iglob < A.i_end) && (jloc < item.jn) && (!item.diag || (jglob > iglob));
          rare = rare && valid;
          off = off || !valid;
        }
        rmask |= (rare ? 1u : 0u) << k;
        double F, gq;
        if constexpr (EXP < 0) distinv_fast<EXP>(u, off, P, F, gq);
        else switching_fast<EXP, PL>(u, off, P, F, gq);
        if constexpr (PL) {
          // F is +0 or positive: test the bit pattern on the integer pipe
          plm[k] |= (__double_as_longlong(F) != 0ll ? 1u : 0u) << t;
        }
        fsum += F;
        a1x[k] = fma(gq, dx, a1x[k]);
        b2x = fma(gq, dx, b2x);
        a1y[k] = fma(gq, dy, a1y[k]);
        b2y = fma(gq, dy, b2y);
        a1z[k] = fma(gq, dz, a1z[k]);
        b2z = fma(gq, dz, b2z);
      }
    }
    if (__builtin_expect(rmask != 0u, 0)) {
      // queue the rare pairs of this step for exact_kernel (cold path)
      do {
        int const k = __ffs(rmask) - 1;
        rmask &= rmask - 1;
        unsigned const slot = atomicAdd(A.rare_count, 1u);
        if (slot < A.rare_cap) {
          A.rare[slot] = ((unsigned long long)(unsigned)(ibase + k * 32) << 32) |
                         (unsigned long long)(unsigned)jglob;
        }
      } while (rmask);
    }
    // the accumulators follow their atom: lane l takes over the atom lane l+1 just did
    int const src = (lane + 1) & 31;
    b2x = __shfl_sync(0xffffffffu, b2x, src);
    b2y = __shfl_sync(0xffffffffu, b2y, src);
    b2z = __shfl_sync(0xffffffffu, b2z, src);
  }
  // lane l now holds the warp's sum for atom tbase + l: add to the block-private copy
  atomicAdd(&ax[tbase + lane], b2x);
  atomicAdd(&ay[tbase + lane], b2y);
  atomicAdd(&az[tbase + lane], b2z);

  if constexpr (PL) {
    // stream compaction of this warp-tile's listed pairs into the global (i,j) list
    int cnt = 0;
#pragma unroll
    for (int k = 0; k < R; k++) cnt += __popc(plm[k]);
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int const v = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += v;
    }
    int const total = __shfl_sync(0xffffffffu, incl, 31);
    if (total > 0) {
      unsigned long long base = 0;
      if (lane == 31) base = atomicAdd(A.pl_count, (unsigned long long)total);
      base = __shfl_sync(0xffffffffu, base, 31);
      unsigned long long off = base + (unsigned long long)(incl - cnt);
#pragma unroll
      for (int k = 0; k < R; k++) {
        unsigned m = plm[k];
        unsigned long long const hi = (unsigned long long)(unsigned)(ibase + k * 32) << 32;
        while (m) {
          int const t = __ffs(m) - 1;
          m &= m - 1;
          if (off < A.pl_cap) {
            A.pl[off] = hi | (unsigned long long)(unsigned)(item.j0 + tbase + ((lane + t) & 31));
          }
          off++;
        }
      }
    }
  }
}

#if defined(B200CV_TRACE)
// debug builds only (make EXTRA=-DB200CV_TRACE): per-CTA timeline, read back by b200cv_trace_dump
__device__ unsigned long long g_trace[6 * 65536];
__device__ __forceinline__ unsigned long long trace_now()
{
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#endif

// ---- the kernel -----------------------------------------------------------------------------
template <int EXP, bool ANISO, int PBC, bool PL, int R, int W, int MINB>
__global__ void __launch_bounds__(W * 32, MINB) sweep_kernel(const __grid_constant__ SweepArgs A)
{
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *sx = reinterpret_cast<double *>(smem_raw);
  double *sy = sx + JCMAX;
  double *sz = sy + JCMAX;
  double *ax = sz + JCMAX;
  double *ay = ax + JCMAX;
  double *az = ay + JCMAX;
  __shared__ __align__(8) uint64_t mbar;
  __shared__ double red[W];
  // general cells: the 13 image shifts (+ a zero row), indexed per lane (general_image)
  __shared__ double tritab[PBC == BC_GENERAL ? 42 : 1];
  if constexpr (PBC == BC_GENERAL) {
    if (threadIdx.x < 42) {
      tritab[threadIdx.x] = threadIdx.x < 39 ? A.P.tri_shift[threadIdx.x / 3][threadIdx.x % 3] : 0.0;
    }
  }

  if (A.gate != nullptr && *A.gate != A.gate_on) {
    if (threadIdx.x == 0) A.partials[blockIdx.x] = 0.0;
    return;
  }
#if defined(B200CV_TRACE)
  unsigned long long const trace_t0 = trace_now();
#endif
  int const item_index = blockIdx.x;
  SweepItem const item = A.items[item_index];
  int const lane = threadIdx.x & 31;
  int const warp = threadIdx.x >> 5;
  int const jn_pad = (item.jn + 31) & ~31;
  int const ntiles = jn_pad >> 5;

  if (threadIdx.x == 0) {
    mbar_init(&mbar, 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t const bytes = (uint32_t)jn_pad * 8u;
    mbar_arrive_expect_tx(&mbar, 3u * bytes);
    bulk_g2s(sx, A.x2 + item.j0, bytes, &mbar);
    bulk_g2s(sy, A.y2 + item.j0, bytes, &mbar);
    bulk_g2s(sz, A.z2 + item.j0, bytes, &mbar);
  }
  for (int t = threadIdx.x; t < jn_pad; t += W * 32) {
    ax[t] = 0.0;
    ay[t] = 0.0;
    az[t] = 0.0;
  }

  // On the specialised non-periodic path rows beyond the group and the padding of the last
  // tile need no masking: they sit at +-1e30 (rows here, padding in the position buffers, see
  // alloc_group), where 1/(1+l2^n') is ~1e-180 and the gradient factor underflows to exactly 0.
  constexpr bool FAR_PADDING = (EXP > 0) && !PL && (PBC == BC_NONE);
  double const row_pad = FAR_PADDING ? 1.0e30 : 0.0;
  // group1 atoms of this lane: i = ibase + 32 k
  int const ibase = item.i0 + warp * (32 * R) + lane;
  double xi[R], yi[R], zi[R], a1x[R], a1y[R], a1z[R];
#pragma unroll
  for (int k = 0; k < R; k++) {
    int const i = ibase + k * 32;
    bool const ok = i < A.i_end;
    xi[k] = ok ? A.x1[i] : row_pad;
    yi[k] = ok ? A.y1[i] : row_pad;
    zi[k] = ok ? A.z1[i] : row_pad;
    a1x[k] = a1y[k] = a1z[k] = 0.0;
  }
  double fsum = 0.0;

  __syncthreads();  // accumulators zeroed
  mbar_wait(&mbar, 0);
#if defined(B200CV_TRACE)
  unsigned long long const trace_t1 = trace_now();
#endif

  // rows of this warp: [iw0, iw1).  Masking (select path) is decided per warp and tile: rows
  // beyond the group, a partial last tile, or -- selfCoordNum -- a tile that reaches into the
  // warp's own rows (j > i has to be applied); tiles entirely at or below the rows are skipped.
  int const iw0 = item.i0 + warp * (32 * R);
  int const iw1 = iw0 + 32 * R;
  bool const masked_rows = !FAR_PADDING && (iw1 > A.i_end);
  int const tile0 = (warp * ntiles) / W;  // staggered start: warps touch different tiles
#pragma unroll 1
  for (int tt = 0; tt < ntiles; tt++) {
    int tile = tile0 + tt;
    if (tile >= ntiles) tile -= ntiles;
    int const jt0 = item.j0 + tile * 32;
    if (item.diag && (jt0 + 32 <= iw0 + 1)) continue;  // no j > i in this tile for this warp
    bool const masked = masked_rows || (!FAR_PADDING && ((tile + 1) * 32 > item.jn)) ||
                        (item.diag && (jt0 < iw1));
    if (masked) {
      sweep_tile<EXP, ANISO, PBC, PL, R, true>(A, item, tile, lane, ibase, sx, sy, sz, ax, ay, az,
                                                xi, yi, zi, a1x, a1y, a1z, fsum, tritab);
    } else {
      sweep_tile<EXP, ANISO, PBC, PL, R, false>(A, item, tile, lane, ibase, sx, sy, sz, ax, ay,
                                                 az, xi, yi, zi, a1x, a1y, a1z, fsum, tritab);
    }
  }

#if defined(B200CV_TRACE)
  unsigned long long const trace_t2 = trace_now();
#endif
  // value: warp tree, then one partial per CTA in a fixed order
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) fsum += __shfl_xor_sync(0xffffffffu, fsum, d);
  if (lane == 0) red[warp] = fsum;
  __syncthreads();  // also: all warps are done adding into ax/ay/az
  if (threadIdx.x == 0) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < W; w++) s += red[w];
    A.partials[item_index] = s;
  }

  if (A.want_grad) {
    // G_d = c_grad[d] * gq * diff_d ; group1 gets -G, group2 +G
    double const cx = A.P.c_grad[0], cy = A.P.c_grad[1], cz = A.P.c_grad[2];
#pragma unroll
    for (int k = 0; k < R; k++) {
      int const i = ibase + k * 32;
      if (i < A.i_end) {
        atomicAdd(&A.g1x[i], -cx * a1x[k]);
        atomicAdd(&A.g1y[i], -cy * a1y[k]);
        atomicAdd(&A.g1z[i], -cz * a1z[k]);
      }
    }
    for (int t = threadIdx.x; t < item.jn; t += W * 32) {
      atomicAdd(&A.g2x[item.j0 + t], cx * ax[t]);
      atomicAdd(&A.g2y[item.j0 + t], cy * ay[t]);
      atomicAdd(&A.g2z[item.j0 + t], cz * az[t]);
    }
  }
#if defined(B200CV_TRACE)
  if (threadIdx.x == 0 && blockIdx.x < 65536) {
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    g_trace[6 * blockIdx.x + 0] = trace_t0;
    g_trace[6 * blockIdx.x + 1] = trace_t1;
    g_trace[6 * blockIdx.x + 2] = trace_t2;
    g_trace[6 * blockIdx.x + 3] = trace_now();
    g_trace[6 * blockIdx.x + 4] = smid;
    g_trace[6 * blockIdx.x + 5] = (unsigned long long)item_index;
  }
#endif
}

constexpr size_t sweep_smem_bytes() { return 6 * JCMAX * sizeof(double); }

}  // namespace b200cv
