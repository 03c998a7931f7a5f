// pair_math.cuh -- per-pair arithmetic of the coordination-number switching function.
//
// Two evaluators:
//
//  * pair_fast<>: the hot path.  FMA-contracted, one reciprocal per pair, integer
//    exponents by repeated multiplication, constants folded out of the loop.  It
//    returns (F, gq) with the pair gradient G_d = C_d * gq * diff_d; the per-axis
//    constants C_d are applied once per atom when accumulators are flushed.
//    It also returns `rare`: the pair lies so close to a decision threshold
//    (|l2 - 1| < 2^-4, where the reference's quotient form and its Taylor branch
//    carry rounding errors of their own that parity has to reproduce; or, with a pair list, l2 ~ l2_max or f0 ~ tolerance) that
//    only arithmetic in the reference's own operation order reproduces the
//    reference.  Rare pairs contribute nothing here; they are queued and
//    evaluated by pair_exact.
//
//  * pair_exact: restates colvar::coordnum::compute_pair_coordnum
//    (src/colvarcomp_coordnums.h:227-281) and switching_function (:168-224) in the
//    reference's operation order with round-to-nearest intrinsics (no FMA
//    contraction), so its l2, F and pair-list decision are bit-identical to the
//    reference CPU path for bit-identical inputs.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200cv {

enum : int { EF_GRADIENTS = 1, EF_USE_PAIRLIST = 1 << 9, EF_REBUILD_PAIRLIST = 1 << 10 };
enum : int { BC_NONE = 0, BC_ORTHO = 1, BC_GENERAL = 2 };

// Everything the pair functions need; lives in kernel parameter (constant) space.
struct PairParams {
  double inv_r0[3];
  double inv_r0sq[3];
  double c_grad[3];  // C_d of the fast path (see pair_fast)
  double tol;        // pairlist tolerance (0 without pair list)
  double inv1mt;     // 1 / (1 - tol)
  double ntol;       // -tol * inv1mt
  double l2_max;     // tolerance_l2_max (1e20 without pair list)
  int en2, ed2;      // en/2, ed/2 (distanceInv: en2 = exponent/2)
  int func;          // 0: switching function (coordination numbers), 1: distanceInv pair term
  // pair-list decision window on the high word of l2: [thr_lo_hi, thr_lo_hi + thr_width] covers
  // l2_max (Newton's approximate root) and the exact root of f0(l2) = tol with one high-word
  // step (2^-20 relative) of margin on both sides
  int thr_lo_hi;
  unsigned thr_width;
  // boundary conditions (src/colvars_system.h)
  int bc_type;       // reference enum: 0 non-periodic, 1 mixed, 2 orthogonal, 3 triclinic
  int periodic[3];
  double cell[3][3];   // unit_cell_x/y/z
  double recip[3][3];  // reciprocal_cell_x/y/z
  // general (mixed / triclinic) cells, image search of the tile sweep (general_image below);
  // filled by b200cv_set_boundaries for bc_type 1 and 3
  double cell2[3][3];        // 2 * unit cell vectors
  double tri_n[13];          // |s_k|^2 of the 13 directions TRI_DIRS (1e300: not periodic)
  double tri_shift[13][3];   // s_k = ix a + iy b + iz c in the reference's operation order
  double tri_margin;         // image decisions closer than this go to the exact-pair queue
  // 0: an orthogonal cell whose vectors carry off-diagonal terms below the reference's tolerance
  // (1e-5; src/colvars_system.h:120-140 keeps type "orthogonal" for it): the reference rounds the
  // fractional coordinates with the full vectors but does not search the neighbouring images
  int tri_search;
  // single-precision copies for the DECISIONS of general_image (which integers, which image):
  // reciprocal vectors, 2 * cell vectors, |s_k|^2 (3e38: not periodic), decision margin
  float recipf[3][3];
  float cell2f[3][3];
  float tri_nf[13];
  float tri_marginf;
};

// the 13 lattice directions of get_triclinic_shift (src/colvars_system.h:194-219) up to sign:
// the 26 neighbouring images are +-s_k
__host__ __device__ constexpr int tri_dir(int k, int axis)
{
  constexpr int D[13][3] = {{1, 0, 0}, {0, 1, 0},  {0, 0, 1},  {1, 1, 0},  {1, -1, 0},
                            {1, 0, 1}, {1, 0, -1}, {0, 1, 1},  {0, 1, -1}, {1, 1, 1},
                            {1, 1, -1}, {1, -1, 1}, {1, -1, -1}};
  return D[k][axis];
}

// ---- reciprocal --------------------------------------------------------------------
// MUFU.RCP64H seed (>= 20 good bits, PTX rcp.approx.ftz.f64) + one cubic
// (Halley-type) step: e = 1 - q*y0; y = y0 + y0*(e + e*e).  3 DFMA, error ~0.5 ulp
// + e^3 (2^-60).  q is always in [1, 1e300] here, so no special-case handling.
__device__ __forceinline__ double rcp_seed(double q)
{
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(q));
  return y;
}

__device__ __forceinline__ double fast_rcp(double q)
{
  double const y0 = rcp_seed(q);
  double const e = fma(-q, y0, 1.0);
#if defined(B200CV_RCP_NEWTON2)
  double const y1 = fma(y0, e, y0);
  double const e1 = fma(-q, y1, 1.0);
  return fma(y1, e1, y1);
#else
  double const t = fma(e, e, e);
  return fma(y0, t, y0);
#endif
}

// x^n for a compile-time n by the shortest multiplication chain the compiler finds
template <int N> __device__ __forceinline__ double ipow(double x)
{
  if constexpr (N == 0) return 1.0;
  else if constexpr (N == 1) return x;
  else if constexpr (N % 2 == 0) { double const h = ipow<N / 2>(x); return h * h; }
  else return x * ipow<N - 1>(x);
}

// runtime integer power, same square-and-multiply order as cvm::integer_power
// (src/colvarmodule.h:105-116) but contracted; used by the runtime-exponent fast path
__device__ __forceinline__ double ipow_rt(double x, int n)
{
  double yy = 1.0, ww = x;
  for (int nn = n; nn != 0; nn >>= 1, ww *= ww) {
    if (nn & 1) yy *= ww;
  }
  return yy;
}

// reciprocal for the general quotient form: the argument b^2*l2 is positive but can be
// tiny (l2 -> 0) or huge; fall back to IEEE division outside the seed's comfortable range.
__device__ __forceinline__ double fast_rcp_wide(double x)
{
  if (x > 1e-290 && x < 1e290) return fast_rcp(x);
  return 1.0 / x;
}

// "|l2 - 1| < 2^-4" on the high word only (integer pipe, not the FP64 pipe).
// Why so wide: the reference's log-derivative m' xd / ((1-xd) l2) - n' xn / ((1-xn) l2)
// subtracts two O(1/h) terms (h = l2 - 1), so ITS rounding error is ~5e-17 / h^2 relative:
// 3e-11 at h = 1e-3, 1e-14 at h = 2^-4.  Any better-conditioned formula disagrees with the
// reference by that much, so parity at 1e-12 needs the reference's own operation order for
// every pair in this window (about 2 pairs per group1 atom at water density -- ~1e-6 of all
// pairs at the benchmark sizes).
__device__ __forceinline__ bool near_one(double u)
{
  return (unsigned)(__double2hiint(u) - 0x3FEE0000) < 0x30000u;
}

// high words within +-1 of each other (relative distance < ~2e-6); both positive
__device__ __forceinline__ bool hi_near(double a, int b_hi)
{
  return (unsigned)(__double2hiint(a) - b_hi + 1) <= 2u;
}

// Is this pair one that only reference-order arithmetic reproduces?  A function of l2 alone,
// evaluated on its high word (integer pipe): |l2 - 1| < 2^-4, or -- with a pair list -- l2
// inside the window that covers l2_max and the root of f0(l2) = tolerance (with ~1e-6
// relative margin), the two places where the reference takes a yes/no decision
// (src/colvarcomp_coordnums.h:203-206,256-261).  Outside that window "l2 > l2_max" and
// "shifted F < 0" are the same statement, and both are far from flipping.
template <bool PL> __device__ __forceinline__ bool pair_is_rare(double u, PairParams const &P)
{
  bool rare = near_one(u);
  if constexpr (PL) rare = rare || ((unsigned)(__double2hiint(u) - P.thr_lo_hi) <= P.thr_width);
  return rare;
}

// l2 value that switches a pair off in the specialised path: F = 2^-300, gq = 2^-400, i.e.
// nothing at any physical scale, and exactly zero with a pair list (l2 > l2_max).
#define B200CV_L2_OFF 1.2676506002282294e30 /* 2^100 */

// EXP = 0: runtime exponents (en2, ed2 from params), general quotient form.
// EXP = n' > 0: compile-time en/2 = n' with ed = 2*en, for which
//   (1 - x^n') / (1 - x^2n') = 1 / (1 + x^n'),  dF/dl2 = -n' x^(n'-1) / (1 + x^n')^2
// (no singularity at l2 = 1, one FMA chain, one reciprocal).
//
// Output convention: G_d = c_grad[d] * gq * diff_d, F >= 0.
//   EXP > 0 : gq = u^(n'-1) r^2,          c_grad[d] = -2 n' inv1mt inv_r0sq[d]
//   EXP = 0 : gq = dF/dl2 (incl. inv1mt), c_grad[d] = 2 inv_r0sq[d]
// `off` (pair is padding, masked, or queued for the exact path) must contribute nothing:
// EXP > 0 replaces l2 by B200CV_L2_OFF up front (one select instead of two),
// EXP = 0 zeroes F and gq at the end (its F does not decay for every exponent choice).
template <int EXP, bool PL>
__device__ __forceinline__ void switching_fast(double u, bool off, PairParams const &P, double &F,
                                               double &gq)
{
  double f0;
  if constexpr (EXP > 0) {
    static_assert(EXP <= 10, "B200CV_L2_OFF^EXP must stay finite");
    u = off ? B200CV_L2_OFF : u;
    double const um1 = ipow<EXP - 1>(u);
    double const q = fma(um1, u, 1.0);
    double const r = fast_rcp(q);
    f0 = r;
    gq = um1 * (r * r);
  } else {
    double const xn = ipow_rt(u, P.en2);
    double const xd = ipow_rt(u, P.ed2);
    double const a = 1.0 - xn;
    double const b = 1.0 - xd;
    // one reciprocal: rinv = 1 / (b^2 l2);  F = a b l2 rinv,
    // dF/dl2 = (m' xd a - n' xn b) rinv
    double const bl = b * u;
    double const rinv = fast_rcp_wide(b * bl);
    double const en2_r = (double)P.en2, ed2_r = (double)P.ed2;
    f0 = (a * bl) * rinv;
    gq = (ed2_r * (xd * a) - en2_r * (xn * b)) * rinv;
    if (u < 1e-280) {  // coincident atoms: F -> 1, dF/dl2 -> 0 (the reference yields NaN here)
      f0 = 1.0;
      gq = 0.0;
    }
  }
  if constexpr (PL) {
    F = fma(f0, P.inv1mt, P.ntol);
    // l2 > l2_max and F < 0 on bit patterns (both sides positive / sign bit): integer pipe
    bool out = (__double_as_longlong(u) > __double_as_longlong(P.l2_max)) ||
               (__double2hiint(F) < 0);
    if constexpr (EXP == 0) {
      gq *= P.inv1mt;
      out = out || off;
    }
    if (out) { F = 0.0; gq = 0.0; }
  } else {
    F = f0;
    if constexpr (EXP == 0) {
      if (off) { F = 0.0; gq = 0.0; }
    }
  }
}

// Specialised form with the pair-list shift but WITHOUT the yes/no decision:
// Fh = (f0 - tol)/(1 - tol) (may be negative), gqh = u^(n'-1) r^2.  The sweep applies the
// decision (Fh > 0) itself and records it, so that a cold block can take a pair out again.
// distanceInv pair term (src/colvarcomp_distances.cpp:415-427): with u = d^2 (inv_r0sq = 1 for
// this function), F = u^(-n/2) and the gradient w.r.t. the second atom is -n F / u * diff, i.e.
// gq = F / u with c_grad = -n.  Well conditioned everywhere: no exact-pair queue needed.
// EXP = -1: runtime n/2 (P.en2);  EXP = -h, h >= 2: compile-time n/2 = h.
template <int EXP>
__device__ __forceinline__ void distinv_fast(double u, bool off, PairParams const &P, double &F,
                                             double &gq)
{
  static_assert(EXP < 0, "distinv_fast is selected by negative EXP");
  u = off ? 1.0 : u;
  // any physical d^2 is far inside the seed's range; d = 0 gives NaN here and 0/0 in the
  // reference's gradient (src/colvarcomp_distances.cpp:419)
  double const r = fast_rcp(u);
  double f;
  if constexpr (EXP == -1) f = ipow_rt(r, P.en2);
  else f = ipow<-EXP>(r);
  F = off ? 0.0 : f;
  gq = off ? 0.0 : f * r;
}

template <int EXP>
__device__ __forceinline__ void switching_shifted(double u, PairParams const &P, double &Fh,
                                                  double &gqh)
{
  static_assert(EXP > 0, "specialised exponents only");
  double const um1 = ipow<EXP - 1>(u);
  double const q = fma(um1, u, 1.0);
  double const r = fast_rcp(q);
  Fh = fma(r, P.inv1mt, P.ntol);
  gqh = um1 * (r * r);
}

// ---- exact (reference-order) path -------------------------------------------------------

// cvm::integer_power (src/colvarmodule.h:105-116), no contraction
__device__ __forceinline__ double integer_power_exact(double x, int n)
{
  if (x == 0.0) return 0.0;
  double yy = 1.0, ww = x;
  for (int nn = n; nn != 0; nn >>= 1, ww = __dmul_rn(ww, ww)) {
    if (nn & 1) yy = __dmul_rn(yy, ww);
  }
  return yy;
}

__device__ __forceinline__ double dot3_exact(const double a[3], double bx, double by, double bz)
{
  return __dadd_rn(__dadd_rn(__dmul_rn(a[0], bx), __dmul_rn(a[1], by)), __dmul_rn(a[2], bz));
}

// system_boundary_conditions::position_distance (src/colvars_system.h:161-219)
__device__ inline void position_distance_exact(PairParams const &P, double x1, double y1,
                                               double z1, double x2, double y2, double z2,
                                               double &dx, double &dy, double &dz)
{
  dx = __dsub_rn(x2, x1);
  dy = __dsub_rn(y2, y1);
  dz = __dsub_rn(z2, z1);
  if (P.bc_type == 0 || P.bc_type == 4) return;
  double const xs = floor(__dadd_rn(dot3_exact(P.recip[0], dx, dy, dz), 0.5));
  double const ys = floor(__dadd_rn(dot3_exact(P.recip[1], dx, dy, dz), 0.5));
  double const zs = floor(__dadd_rn(dot3_exact(P.recip[2], dx, dy, dz), 0.5));
  dx = __dsub_rn(dx, __dadd_rn(__dadd_rn(__dmul_rn(xs, P.cell[0][0]), __dmul_rn(ys, P.cell[1][0])),
                               __dmul_rn(zs, P.cell[2][0])));
  dy = __dsub_rn(dy, __dadd_rn(__dadd_rn(__dmul_rn(xs, P.cell[0][1]), __dmul_rn(ys, P.cell[1][1])),
                               __dmul_rn(zs, P.cell[2][1])));
  dz = __dsub_rn(dz, __dadd_rn(__dadd_rn(__dmul_rn(xs, P.cell[0][2]), __dmul_rn(ys, P.cell[1][2])),
                               __dmul_rn(zs, P.cell[2][2])));
  if (P.bc_type != 2 && P.tri_search) {
    // get_triclinic_shift: search the neighbouring images for a shorter vector
    double min_d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
    double rx = 0.0, ry = 0.0, rz = 0.0;
    int const nx = P.periodic[0] ? 1 : 0, ny = P.periodic[1] ? 1 : 0, nz = P.periodic[2] ? 1 : 0;
    for (int ix = -nx; ix <= nx; ix++) {
      for (int iy = -ny; iy <= ny; iy++) {
        for (int iz = -nz; iz <= nz; iz++) {
          double const sx = __dadd_rn(__dadd_rn(__dmul_rn((double)ix, P.cell[0][0]),
                                                __dmul_rn((double)iy, P.cell[1][0])),
                                      __dmul_rn((double)iz, P.cell[2][0]));
          double const sy = __dadd_rn(__dadd_rn(__dmul_rn((double)ix, P.cell[0][1]),
                                                __dmul_rn((double)iy, P.cell[1][1])),
                                      __dmul_rn((double)iz, P.cell[2][1]));
          double const sz = __dadd_rn(__dadd_rn(__dmul_rn((double)ix, P.cell[0][2]),
                                                __dmul_rn((double)iy, P.cell[1][2])),
                                      __dmul_rn((double)iz, P.cell[2][2]));
          double const tx = __dadd_rn(dx, sx), ty = __dadd_rn(dy, sy), tz = __dadd_rn(dz, sz);
          double const d2 =
              __dadd_rn(__dadd_rn(__dmul_rn(tx, tx), __dmul_rn(ty, ty)), __dmul_rn(tz, tz));
          if (d2 < min_d2) { rx = sx; ry = sy; rz = sz; min_d2 = d2; }
        }
      }
    }
    dx = __dadd_rn(dx, rx);
    dy = __dadd_rn(dy, ry);
    dz = __dadd_rn(dz, rz);
  }
}

// position_distance for mixed / triclinic cells on the tile sweep.
//
// The reference (src/colvars_system.h:161-219) takes two kinds of decisions per pair -- the three
// integers of the fractional rounding (steps 1-2) and which of the 27 neighbouring images is the
// shortest (step 3, get_triclinic_shift: zero shift and then loop order winning ties) -- and
// otherwise does a handful of double-precision operations with them.  Here the DECISIONS are
// taken in single precision (the FP32 pipe runs next to the FP64 pipe and at eight times its
// rate), each with a margin several times the worst single-precision error; the arithmetic
// with the decided integers / image is then done in double precision in the reference's
// operation order, so the image vector is bit-identical whenever the decisions are.  A decision
// inside its margin is not taken: the function returns true and the pair goes to the exact-pair
// queue, where pair_exact runs the reference's loop (4 pairs in 10^4 in a rhombic dodecahedron;
// taking the decision again in double precision in the loop -- a call, for the lanes concerned --
// was measured and costs more than it saves: the warp waits for it once in 30 pairs).
//   steps 1-2  u = recip_a . d in single precision, error below 2^-21.5 * sum |terms|; floor(u + 1/2)
//              is the integer next to u unless u is within the margin, 2^-19 * sum |terms|, of a
//              half-integer
//   step 3     |d +- s_k|^2 - |d|^2 = |s_k|^2 +- 2 d.s_k ,  min over the sign = |s_k|^2 - |2 d.s_k|
//              for the 13 directions s_k (three dot products, ten sums, thirteen differences, a
//              running minimum that carries k in the four lowest mantissa bits).  Error of a
//              score (rounding + the tag) below 2^-17.8 of the largest |s_k|^2; margin 2^-15 of it:
//              another candidate -- or the zero shift -- within the margin of the best one makes
//              the pair ambiguous.
// The chosen shift is added as the reference adds it (tab[k] holds s_k in the reference's
// operation order; the sign flip is exact).
// tab: 14 x 3 doubles in shared memory, rows 0..12 = P.tri_shift, row 13 = 0.
// (B200CV_TRI_FP64: double precision only.)
// every decision in double precision (margin 2^-26 of the largest |s_k|^2)
__device__ __forceinline__ bool general_image_f64(PairParams const &P, const double *tab, double &dx,
                                                  double &dy, double &dz)
{

  double const xs = floor(__dadd_rn(dot3_exact(P.recip[0], dx, dy, dz), 0.5));
  double const ys = floor(__dadd_rn(dot3_exact(P.recip[1], dx, dy, dz), 0.5));
  double const zs = floor(__dadd_rn(dot3_exact(P.recip[2], dx, dy, dz), 0.5));
  dx = __dsub_rn(dx, __dadd_rn(__dadd_rn(__dmul_rn(xs, P.cell[0][0]), __dmul_rn(ys, P.cell[1][0])),
                               __dmul_rn(zs, P.cell[2][0])));
  dy = __dsub_rn(dy, __dadd_rn(__dadd_rn(__dmul_rn(xs, P.cell[0][1]), __dmul_rn(ys, P.cell[1][1])),
                               __dmul_rn(zs, P.cell[2][1])));
  dz = __dsub_rn(dz, __dadd_rn(__dadd_rn(__dmul_rn(xs, P.cell[0][2]), __dmul_rn(ys, P.cell[1][2])),
                               __dmul_rn(zs, P.cell[2][2])));
  double const pa = fma(dz, P.cell2[0][2], fma(dy, P.cell2[0][1], dx * P.cell2[0][0]));
  double const pb = fma(dz, P.cell2[1][2], fma(dy, P.cell2[1][1], dx * P.cell2[1][0]));
  double const pc = fma(dz, P.cell2[2][2], fma(dy, P.cell2[2][1], dx * P.cell2[2][0]));
  double const pab = pa + pb, pamb = pa - pb;
  double const q[13] = {pa,      pb,      pc,       pab,      pamb,      pa + pc,  pa - pc,
                        pb + pc, pb - pc, pab + pc, pab - pc, pamb + pc, pamb - pc};
  double m[13];
#pragma unroll
  for (int k = 0; k < 13; k++) {
    double const v = P.tri_n[k] - fabs(q[k]);
    m[k] = __hiloint2double(__double2hiint(v), (__double2loint(v) & ~15) | k);
  }
  // minimum as a tree (depth 4): the search of one pair is a single dependency chain otherwise
  double const b0 = fmin(m[0], m[1]), b1 = fmin(m[2], m[3]), b2 = fmin(m[4], m[5]);
  double const b3 = fmin(m[6], m[7]), b4 = fmin(m[8], m[9]), b5 = fmin(m[10], m[11]);
  double const c0 = fmin(b0, b1), c1 = fmin(b2, b3), c2 = fmin(b4, b5);
  double const best = fmin(fmin(c0, c1), fmin(c2, m[12]));
  double const thr = best + P.tri_margin;
  int cnt = 0;
#pragma unroll
  for (int k = 0; k < 13; k++) cnt += (m[k] < thr) ? 1 : 0;
  bool const shifted = best < P.tri_margin;
  bool const ambiguous = shifted && (cnt >= 2 || best > -P.tri_margin);
  int const idx = shifted ? (__double2loint(best) & 15) : 13;
  double const sx = tab[3 * idx], sy = tab[3 * idx + 1], sz = tab[3 * idx + 2];
  double const t = fma(dz, sz, fma(dy, sy, dx * sx));
  double const sgn = t > 0.0 ? -1.0 : 1.0;
  dx = fma(sgn, sx, dx);
  dy = fma(sgn, sy, dy);
  dz = fma(sgn, sz, dz);
  return ambiguous;
}

// the decisions in single precision
__device__ __forceinline__ bool general_image_f32(PairParams const &P, const double *tab, double &dx,
                                                  double &dy, double &dz)
{
  bool amb = false;
  double sh[3];
  {
    float const fx = (float)dx, fy = (float)dy, fz = (float)dz;
#pragma unroll
    for (int a = 0; a < 3; a++) {
      // floor(u + 1/2) is the integer next to u unless u is next to a half-integer
      float const t0 = P.recipf[a][0] * fx, t1 = P.recipf[a][1] * fy, t2 = P.recipf[a][2] * fz;
      float const u = (t0 + t1) + t2;
      float const mag = (fabsf(t0) + fabsf(t1)) + fabsf(t2);
      float const r = rintf(u);
      float const dev = fabsf(u - r);  // in [0, 1/2]
      amb = amb || !(fmaf(mag, 0x1p-19f, dev) < 0.5f - 0x1p-30f);  // (a NaN is ambiguous)
      sh[a] = (double)r;
    }
  }
  dx = __dsub_rn(dx, __dadd_rn(__dadd_rn(__dmul_rn(sh[0], P.cell[0][0]), __dmul_rn(sh[1], P.cell[1][0])),
                               __dmul_rn(sh[2], P.cell[2][0])));
  dy = __dsub_rn(dy, __dadd_rn(__dadd_rn(__dmul_rn(sh[0], P.cell[0][1]), __dmul_rn(sh[1], P.cell[1][1])),
                               __dmul_rn(sh[2], P.cell[2][1])));
  dz = __dsub_rn(dz, __dadd_rn(__dadd_rn(__dmul_rn(sh[0], P.cell[0][2]), __dmul_rn(sh[1], P.cell[1][2])),
                               __dmul_rn(sh[2], P.cell[2][2])));
  float const gx = (float)dx, gy = (float)dy, gz = (float)dz;
  float const pa = fmaf(gz, P.cell2f[0][2], fmaf(gy, P.cell2f[0][1], gx * P.cell2f[0][0]));
  float const pb = fmaf(gz, P.cell2f[1][2], fmaf(gy, P.cell2f[1][1], gx * P.cell2f[1][0]));
  float const pc = fmaf(gz, P.cell2f[2][2], fmaf(gy, P.cell2f[2][1], gx * P.cell2f[2][0]));
  float const pab = pa + pb, pamb = pa - pb;
  float const q[13] = {pa,      pb,      pc,       pab,      pamb,      pa + pc,  pa - pc,
                       pb + pc, pb - pc, pab + pc, pab - pc, pamb + pc, pamb - pc};
  float m[13];
#pragma unroll
  for (int k = 0; k < 13; k++) {
    float const v = P.tri_nf[k] - fabsf(q[k]);
    m[k] = __int_as_float((__float_as_int(v) & ~15) | k);
  }
  // minimum as a tree (depth 4): the search of one pair is a single dependency chain otherwise
  float const b0 = fminf(m[0], m[1]), b1 = fminf(m[2], m[3]), b2 = fminf(m[4], m[5]);
  float const b3 = fminf(m[6], m[7]), b4 = fminf(m[8], m[9]), b5 = fminf(m[10], m[11]);
  float const c0 = fminf(b0, b1), c1 = fminf(b2, b3), c2 = fminf(b4, b5);
  float const best = fminf(fminf(c0, c1), fminf(c2, m[12]));
  float const thr = best + P.tri_marginf;
  int cnt = 0;
#pragma unroll
  for (int k = 0; k < 13; k++) cnt += (m[k] < thr) ? 1 : 0;
  bool const shifted = best < P.tri_marginf;
  amb = amb || (shifted && (cnt >= 2 || best > -P.tri_marginf)) || !(best == best);
  int const idx = shifted ? (__float_as_int(best) & 15) : 13;
  double const sx = tab[3 * idx], sy = tab[3 * idx + 1], sz = tab[3 * idx + 2];
  double const t = fma(dz, sz, fma(dy, sy, dx * sx));
  double const sgn = t > 0.0 ? -1.0 : 1.0;
  dx = fma(sgn, sx, dx);
  dy = fma(sgn, sy, dy);
  dz = fma(sgn, sz, dz);
  return amb;
}

__device__ __forceinline__ bool general_image(PairParams const &P, const double *tab, double &dx,
                                              double &dy, double &dz)
{
#if defined(B200CV_TRI_FP64)
  return general_image_f64(P, tab, dx, dy, dz);
#else
  return general_image_f32(P, tab, dx, dy, dz);
#endif
}


// switching_function<flags> (src/colvarcomp_coordnums.h:168-224)
__device__ inline double switching_exact(double l2, double &dFdl2, PairParams const &P,
                                         int flags)
{
  double const xn = integer_power_exact(l2, P.en2);
  double const xd = integer_power_exact(l2, P.ed2);
  double const eps_l2 = 1.0e-7;
  double const h = __dsub_rn(l2, 1.0);
  double const en2_r = (double)P.en2, ed2_r = (double)P.ed2;
  bool const taylor = fabs(h) < eps_l2;
  double f0;
  if (taylor) {
    double const c0 = __ddiv_rn(en2_r, ed2_r);
    double const c1 = __ddiv_rn(__dmul_rn(en2_r, __dsub_rn(en2_r, ed2_r)), __dmul_rn(2.0, ed2_r));
    double const c2 =
        __ddiv_rn(__dmul_rn(__dmul_rn(en2_r, __dsub_rn(en2_r, ed2_r)),
                            __dsub_rn(__dsub_rn(__dmul_rn(2.0, en2_r), ed2_r), 3.0)),
                  __dmul_rn(12.0, ed2_r));
    f0 = __dadd_rn(c0, __dmul_rn(h, __dadd_rn(c1, __dmul_rn(h, c2))));
  } else {
    f0 = __ddiv_rn(__dsub_rn(1.0, xn), __dsub_rn(1.0, xd));
  }
  double func, inv1mt = 0.0;
  if (flags & EF_USE_PAIRLIST) {
    inv1mt = __ddiv_rn(1.0, __dsub_rn(1.0, P.tol));
    func = __dmul_rn(__dsub_rn(f0, P.tol), inv1mt);
  } else {
    func = f0;
  }
  if (func < 0) return 0.0;
  if (flags & EF_GRADIENTS) {
    double log_deriv;
    if (taylor) {
      double const g0 = __dmul_rn(0.5, __dsub_rn(en2_r, ed2_r));
      double const g1 = __ddiv_rn(
          __dmul_rn(__dsub_rn(en2_r, ed2_r), __dsub_rn(__dadd_rn(en2_r, ed2_r), 6.0)), 12.0);
      log_deriv = __dadd_rn(g0, __dmul_rn(h, g1));
    } else {
      double const t1 = __ddiv_rn(__dmul_rn(ed2_r, xd), __dmul_rn(__dsub_rn(1.0, xd), l2));
      double const t2 = __ddiv_rn(__dmul_rn(en2_r, xn), __dmul_rn(__dsub_rn(1.0, xn), l2));
      log_deriv = __dsub_rn(t1, t2);
    }
    dFdl2 = (flags & EF_USE_PAIRLIST) ? __dmul_rn(__dmul_rn(f0, inv1mt), log_deriv)
                                      : __dmul_rn(func, log_deriv);
  }
  return func;
}

// compute_pair_coordnum<flags> (src/colvarcomp_coordnums.h:227-281); returns F and the
// gradient vector G (group1 gets -G, group2 +G); G = 0 when F == 0 or gradients are off
__device__ inline double pair_exact(PairParams const &P, double x1, double y1, double z1,
                                    double x2, double y2, double z2, int flags, double &Gx,
                                    double &Gy, double &Gz)
{
  double dx, dy, dz;
  position_distance_exact(P, x1, y1, z1, x2, y2, z2, dx, dy, dz);
  if (P.func == 1) {
    // distance_inv::calc_value, reference operation order (src/colvarcomp_distances.cpp:415-427)
    double const d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
    double const pw = integer_power_exact(d2, P.en2);
    double const dinv = (d2 == 0.0) ? 0.0 : __ddiv_rn(1.0, pw);
    double const c = __dmul_rn(__ddiv_rn(__dmul_rn((double)(-P.en2), dinv), d2), 2.0);
    Gx = __dmul_rn(c, dx);
    Gy = __dmul_rn(c, dy);
    Gz = __dmul_rn(c, dz);
    return dinv;
  }
  double const sx = __dmul_rn(dx, P.inv_r0[0]);
  double const sy = __dmul_rn(dy, P.inv_r0[1]);
  double const sz = __dmul_rn(dz, P.inv_r0[2]);
  double const l2 = __dadd_rn(__dadd_rn(__dmul_rn(sx, sx), __dmul_rn(sy, sy)), __dmul_rn(sz, sz));
  Gx = Gy = Gz = 0.0;
  if (flags & EF_USE_PAIRLIST) {
    if (l2 > P.l2_max) return 0.0;
  }
  double dFdl2 = 0.0;
  double const F = switching_exact(l2, dFdl2, P, flags);
  if ((flags & EF_GRADIENTS) && (F > 0.0)) {
    Gx = __dmul_rn(dFdl2, __dmul_rn(__dmul_rn(2.0, P.inv_r0sq[0]), dx));
    Gy = __dmul_rn(dFdl2, __dmul_rn(__dmul_rn(2.0, P.inv_r0sq[1]), dy));
    Gz = __dmul_rn(dFdl2, __dmul_rn(__dmul_rn(2.0, P.inv_r0sq[2]), dz));
  }
  return F;
}

}  // namespace b200cv
