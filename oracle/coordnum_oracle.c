/*
 * coordnum_oracle.c -- TEST INFRASTRUCTURE ONLY (see coordnum_oracle.h).
 *
 * CPU restatement of src/colvarcomp_coordnums.{h,cpp}, the boundary-condition
 * helper of src/colvars_system.h and the atom-group loops of
 * src/colvaratoms.cpp that the coordination-number path touches.
 * Build: gcc -O2 -ffp-contract=off (no FMA contraction, IEEE operation order).
 */
#include "coordnum_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* src/colvarmodule.h:105-116: square-and-multiply, x == 0 -> 0 */
double cvo_integer_power(double x, int n)
{
  double yy, ww;
  if (x == 0.0) return 0.0;
  int nn = (n > 0) ? n : -n;
  ww = x;
  for (yy = 1.0; nn != 0; nn >>= 1, ww *= ww) {
    if (nn & 1) yy *= ww;
  }
  return (n > 0) ? yy : 1.0 / yy;
}

/* src/colvarcomp_coordnums.h:168-224 */
double cvo_switching_function(double l2, double *dFdl2, int en, int ed, double pairlist_tol,
                              int flags)
{
  int const en2 = en / 2;
  int const ed2 = ed / 2;

  double const xn = cvo_integer_power(l2, en2);
  double const xd = cvo_integer_power(l2, ed2);
  double const eps_l2 = 1.0e-7;
  double const h = l2 - 1.0;
  double const en2_r = (double)en2;
  double const ed2_r = (double)ed2;
  double func_no_pairlist;

  if (fabs(h) < eps_l2) {
    /* order-2 Taylor expansion around l2 = 1 (:186-190) */
    double const c0 = en2_r / ed2_r;
    double const c1 = (en2_r * (en2_r - ed2_r)) / (2.0 * ed2_r);
    double const c2 = (en2_r * (en2_r - ed2_r) * (2.0 * en2_r - ed2_r - 3.0)) / (12.0 * ed2_r);
    func_no_pairlist = c0 + h * (c1 + h * c2);
  } else {
    func_no_pairlist = (1.0 - xn) / (1.0 - xd);
  }

  double func, inv_one_pairlist_tol = 0.0;
  if (flags & CVO_EF_USE_PAIRLIST) {
    inv_one_pairlist_tol = 1 / (1.0 - pairlist_tol);
    func = (func_no_pairlist - pairlist_tol) * inv_one_pairlist_tol;
  } else {
    func = func_no_pairlist;
  }

  /* negative after the tolerance shift: excluded, no gradient (:203-206) */
  if (func < 0) return 0;

  if (flags & CVO_EF_GRADIENTS) {
    double log_deriv;
    if (fabs(h) < eps_l2) {
      double const g0 = 0.5 * (en2_r - ed2_r);
      double const g1 = ((en2_r - ed2_r) * (en2_r + ed2_r - 6.0)) / 12.0;
      log_deriv = g0 + h * g1;
    } else {
      log_deriv = (ed2_r * xd / ((1.0 - xd) * l2)) - (en2_r * xn / ((1.0 - xn) * l2));
    }
    *dFdl2 = (flags & CVO_EF_USE_PAIRLIST) ? func_no_pairlist * inv_one_pairlist_tol * log_deriv
                                           : func * log_deriv;
  }
  return func;
}

/* src/colvarcomp_coordnums.cpp:162-184: Newton from l2 = 1.001 on F_pairlist(l2) = 0 */
double cvo_compute_tolerance_l2_max(int en, int ed, double tolerance)
{
  double l2 = 1.001;
  double F = 0.0;
  double dFdl2 = 0.0;
  const size_t num_iters_max = 1000000;
  const double result_tol = 1.0e-6;
  const double dF_tol = 1.0e-9;
  size_t i;
  for (i = 0; i < num_iters_max; i++) {
    F = cvo_switching_function(l2, &dFdl2, en, ed, tolerance,
                               CVO_EF_USE_PAIRLIST | CVO_EF_GRADIENTS);
    if ((fabs(F) < result_tol) || (fabs(dFdl2) < dF_tol)) {
      break;
    }
    l2 -= F / dFdl2;
  }
  return l2;
}

/* src/colvarcomp_coordnums.cpp:28-43 */
void cvo_update_cutoffs(cvo_params *p, double r0x, double r0y, double r0z)
{
  p->inv_r0[0] = 1.0 / r0x;
  p->inv_r0[1] = 1.0 / r0y;
  p->inv_r0[2] = 1.0 / r0z;
  p->inv_r0sq[0] = p->inv_r0[0] * p->inv_r0[0];
  p->inv_r0sq[1] = p->inv_r0[1] * p->inv_r0[1];
  p->inv_r0sq[2] = p->inv_r0[2] * p->inv_r0[2];
}

/* ---- boundary conditions ----------------------------------------------------- */

static double dot3(const double a[3], const double b[3])
{
  return a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
}

static void cross3(const double a[3], const double b[3], double c[3])
{
  /* cvm::rvector::outer, src/colvartypes.h */
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = -a[0] * b[2] + a[2] * b[0];
  c[2] = a[0] * b[1] - a[1] * b[0];
}

/* src/colvars_system.h:48-58 */
void cvo_boundaries_reset(cvo_boundaries *bc)
{
  memset(bc, 0, sizeof(*bc));
  bc->type = CVO_BC_NON_PERIODIC;
}

/* src/colvars_system.h:80-158 */
void cvo_set_boundaries(cvo_boundaries *bc, int px, int py, int pz, const double A[3],
                        const double B[3], const double C[3])
{
  const double diagonal_tol2 = 1.0e-10;
  int d;
  bc->periodic[0] = px;
  bc->periodic[1] = py;
  bc->periodic[2] = pz;
  memset(bc->reciprocal_cell, 0, sizeof(bc->reciprocal_cell));

  if ((px == py) && (px == pz)) {
    if (px) {
      bc->type = CVO_BC_PBC_ORTHOGONAL;
    } else {
      cvo_boundaries_reset(bc);
      return;
    }
  } else {
    bc->type = CVO_BC_MIXED;
  }

  int off_diagonal = 0;
  const double *V[3] = {A, B, C};
  for (d = 0; d < 3; d++) {
    if (bc->periodic[d]) {
      int const o1 = (d + 1) % 3, o2 = (d + 2) % 3;
      memcpy(bc->unit_cell[d], V[d], 3 * sizeof(double));
      if ((V[d][o1] * V[d][o1]) > diagonal_tol2 || (V[d][o2] * V[d][o2]) > diagonal_tol2) {
        off_diagonal = 1;
      } else {
        double const n2 = dot3(V[d], V[d]);
        bc->reciprocal_cell[d][0] = V[d][0] / n2;
        bc->reciprocal_cell[d][1] = V[d][1] / n2;
        bc->reciprocal_cell[d][2] = V[d][2] / n2;
      }
    } else {
      memset(bc->unit_cell[d], 0, 3 * sizeof(double));
      memset(bc->reciprocal_cell[d], 0, 3 * sizeof(double));
    }
  }

  if (bc->type == CVO_BC_PBC_ORTHOGONAL && off_diagonal) {
    bc->type = CVO_BC_PBC_TRICLINIC;
  }

  if (bc->type == CVO_BC_PBC_TRICLINIC) {
    double v[3], s;
    cross3(bc->unit_cell[1], bc->unit_cell[2], v);
    s = dot3(v, bc->unit_cell[0]);
    for (d = 0; d < 3; d++) bc->reciprocal_cell[0][d] = v[d] / s;
    cross3(bc->unit_cell[2], bc->unit_cell[0], v);
    s = dot3(v, bc->unit_cell[1]);
    for (d = 0; d < 3; d++) bc->reciprocal_cell[1][d] = v[d] / s;
    cross3(bc->unit_cell[0], bc->unit_cell[1], v);
    s = dot3(v, bc->unit_cell[2]);
    for (d = 0; d < 3; d++) bc->reciprocal_cell[2][d] = v[d] / s;
  }
}

/* src/colvars_system.h:194-219 */
static void get_triclinic_shift(const cvo_boundaries *bc, const double diff[3], double result[3])
{
  double min_dist2 = dot3(diff, diff);
  result[0] = result[1] = result[2] = 0.0;
  int const nx = bc->periodic[0] ? 1 : 0;
  int const ny = bc->periodic[1] ? 1 : 0;
  int const nz = bc->periodic[2] ? 1 : 0;
  for (int ix = -nx; ix <= nx; ix++) {
    for (int iy = -ny; iy <= ny; iy++) {
      for (int iz = -nz; iz <= nz; iz++) {
        double shift[3], t[3];
        for (int d = 0; d < 3; d++) {
          shift[d] = ix * bc->unit_cell[0][d] + iy * bc->unit_cell[1][d] + iz * bc->unit_cell[2][d];
          t[d] = diff[d] + shift[d];
        }
        double const this_dist2 = dot3(t, t);
        if (this_dist2 < min_dist2) {
          result[0] = shift[0];
          result[1] = shift[1];
          result[2] = shift[2];
          min_dist2 = this_dist2;
        }
      }
    }
  }
}

/* src/colvars_system.h:161-191 */
void cvo_position_distance(const cvo_boundaries *bc, const double pos1[3], const double pos2[3],
                           double diff[3])
{
  diff[0] = pos2[0] - pos1[0];
  diff[1] = pos2[1] - pos1[1];
  diff[2] = pos2[2] - pos1[2];

  if (bc->type == CVO_BC_NON_PERIODIC || bc->type == CVO_BC_UNSUPPORTED) {
    return;
  }

  double const x_shift = floor(dot3(bc->reciprocal_cell[0], diff) + 0.5);
  double const y_shift = floor(dot3(bc->reciprocal_cell[1], diff) + 0.5);
  double const z_shift = floor(dot3(bc->reciprocal_cell[2], diff) + 0.5);

  diff[0] -= x_shift * bc->unit_cell[0][0] + y_shift * bc->unit_cell[1][0] +
             z_shift * bc->unit_cell[2][0];
  diff[1] -= x_shift * bc->unit_cell[0][1] + y_shift * bc->unit_cell[1][1] +
             z_shift * bc->unit_cell[2][1];
  diff[2] -= x_shift * bc->unit_cell[0][2] + y_shift * bc->unit_cell[1][2] +
             z_shift * bc->unit_cell[2][2];

  if (bc->type != CVO_BC_PBC_ORTHOGONAL) {
    double s[3];
    get_triclinic_shift(bc, diff, s);
    diff[0] += s[0];
    diff[1] += s[1];
    diff[2] += s[2];
  }
}

/* ---- pair kernel ------------------------------------------------------------- */

/* src/colvarcomp_coordnums.h:227-281 */
double cvo_compute_pair_coordnum(const cvo_params *p, const double a1[3], const double a2[3],
                                 double g1[3], double g2[3], int flags)
{
  double diff[3];
  cvo_position_distance(&p->bc, a1, a2, diff);
  double const sx = diff[0] * p->inv_r0[0];
  double const sy = diff[1] * p->inv_r0[1];
  double const sz = diff[2] * p->inv_r0[2];
  double const l2 = sx * sx + sy * sy + sz * sz; /* rvector::norm2 */
  if (flags & CVO_EF_USE_PAIRLIST) {
    if (l2 > p->tolerance_l2_max) {
      return 0.0;
    }
  }

  double dFdl2 = 0.0;
  double const F = cvo_switching_function(l2, &dFdl2, p->en, p->ed, p->tolerance, flags);

  if ((flags & CVO_EF_GRADIENTS) && (F > 0.0)) {
    double const dl2dx = (2.0 * p->inv_r0sq[0]) * diff[0];
    double const dl2dy = (2.0 * p->inv_r0sq[1]) * diff[1];
    double const dl2dz = (2.0 * p->inv_r0sq[2]) * diff[2];
    double const Gx = dFdl2 * dl2dx, Gy = dFdl2 * dl2dy, Gz = dFdl2 * dl2dz;
    g1[0] += -1.0 * Gx;
    g1[1] += -1.0 * Gy;
    g1[2] += -1.0 * Gz;
    g2[0] += Gx;
    g2[1] += Gy;
    g2[2] += Gz;
  }
  return F;
}

/* ---- atom-group loops --------------------------------------------------------- */

/* src/colvaratoms.cpp:1415-1431: assignment, not accumulation */
void cvo_set_weighted_gradient(size_t n, const double *weight, const double g[3],
                               double *grad_soa)
{
  for (size_t i = 0; i < n; i++) {
    grad_soa[i] = weight[i] * g[0];
    grad_soa[n + i] = weight[i] * g[1];
    grad_soa[2 * n + i] = weight[i] * g[2];
  }
}

/* src/colvarcomp_coordnums.cpp:187-248 */
double cvo_coordnum(const cvo_params *p, size_t n1, const double *pos1, size_t n2,
                    const double *pos2, int use_g1_com, int use_g2_com, const double com1[3],
                    const double com2[3], const double *weight1, const double *weight2,
                    double *grad1, double *grad2, unsigned char *pairlist, int flags)
{
  size_t const n1c = use_g1_com ? 1 : n1;
  size_t const n2c = use_g2_com ? 1 : n2;
  double g1_com_grad[3] = {0, 0, 0}, g2_com_grad[3] = {0, 0, 0};
  unsigned char *pl = pairlist;
  double value = 0.0;

  for (size_t i = 0; i < n1c; ++i) {
    double a1[3], g1[3];
    if (use_g1_com) {
      memcpy(a1, com1, sizeof(a1));
      memcpy(g1, g1_com_grad, sizeof(g1));
    } else {
      a1[0] = pos1[i];
      a1[1] = pos1[n1 + i];
      a1[2] = pos1[2 * n1 + i];
      g1[0] = grad1[i];
      g1[1] = grad1[n1 + i];
      g1[2] = grad1[2 * n1 + i];
    }
    for (size_t j = 0; j < n2c; ++j) {
      double a2[3], g2[3];
      if (use_g2_com) {
        memcpy(a2, com2, sizeof(a2));
        memcpy(g2, g2_com_grad, sizeof(g2));
      } else {
        a2[0] = pos2[j];
        a2[1] = pos2[n2 + j];
        a2[2] = pos2[2 * n2 + j];
        g2[0] = grad2[j];
        g2[1] = grad2[n2 + j];
        g2[2] = grad2[2 * n2 + j];
      }
      int const within = ((flags & CVO_EF_USE_PAIRLIST) &&
                          (*pl || (flags & CVO_EF_REBUILD_PAIRLIST))) ||
                         !(flags & CVO_EF_USE_PAIRLIST);
      double const partial = within ? cvo_compute_pair_coordnum(p, a1, a2, g1, g2, flags) : 0.0;
      if ((flags & CVO_EF_USE_PAIRLIST) && (flags & CVO_EF_REBUILD_PAIRLIST)) {
        *pl = partial > 0.0 ? 1 : 0;
      }
      value += partial;
      if (flags & CVO_EF_USE_PAIRLIST) pl++;
      if (use_g2_com) {
        memcpy(g2_com_grad, g2, sizeof(g2));
      } else {
        grad2[j] = g2[0];
        grad2[n2 + j] = g2[1];
        grad2[2 * n2 + j] = g2[2];
      }
    }
    if (use_g1_com) {
      memcpy(g1_com_grad, g1, sizeof(g1));
    } else {
      grad1[i] = g1[0];
      grad1[n1 + i] = g1[1];
      grad1[2 * n1 + i] = g1[2];
    }
  }

  /* only meaningful when gradients were requested; the reference calls these
   * unconditionally (:242-247), assigning zeros otherwise */
  if (use_g1_com) cvo_set_weighted_gradient(n1, weight1, g1_com_grad, grad1);
  if (use_g2_com) cvo_set_weighted_gradient(n2, weight2, g2_com_grad, grad2);
  return value;
}

/* src/colvarcomp_coordnums.cpp:478-523 */
double cvo_selfcoordnum(const cvo_params *p, size_t n, const double *pos, double *grad,
                        unsigned char *pairlist, int flags)
{
  unsigned char *pl = pairlist;
  double value = 0.0;
  if (n < 2) return 0.0;
  for (size_t i = 0; i < n - 1; i++) {
    double const a1[3] = {pos[i], pos[n + i], pos[2 * n + i]};
    double g1[3] = {grad[i], grad[n + i], grad[2 * n + i]};
    for (size_t j = i + 1; j < n; j++) {
      double const a2[3] = {pos[j], pos[n + j], pos[2 * n + j]};
      double g2[3] = {grad[j], grad[n + j], grad[2 * n + j]};
      int const within = ((flags & CVO_EF_USE_PAIRLIST) &&
                          (*pl || (flags & CVO_EF_REBUILD_PAIRLIST))) ||
                         !(flags & CVO_EF_USE_PAIRLIST);
      double const partial = within ? cvo_compute_pair_coordnum(p, a1, a2, g1, g2, flags) : 0.0;
      if ((flags & CVO_EF_USE_PAIRLIST) && (flags & CVO_EF_REBUILD_PAIRLIST)) {
        *pl = partial > 0.0 ? 1 : 0;
      }
      value += partial;
      if (flags & CVO_EF_USE_PAIRLIST) pl++;
      grad[j] = g2[0];
      grad[n + j] = g2[1];
      grad[2 * n + j] = g2[2];
    }
    grad[i] = g1[0];
    grad[n + i] = g1[1];
    grad[2 * n + i] = g1[2];
  }
  return value;
}

/* ---- OpenMP checkers (row-parallel; summation order differs from the reference) ----------
 * Same pair function as the sequential loops above.  The scalar is accumulated in extended
 * precision (long double, 64-bit mantissa on x86-64): at the BASELINE sizes the sum has up to
 * 1e11 positive terms spanning 12 orders of magnitude, and a plain double running sum -- the
 * reference's own `x.real_value += partial` included -- drops the part of every term that lies
 * below half an ulp of the running sum (measured at 100 k x 1 M: 2.5e-10 relative between a
 * 16-thread double sum and this one).  These checkers therefore return the correctly rounded sum
 * of the reference's per-pair values; the sequential functions above stay in the reference's
 * order and precision. */

double cvo_coordnum_omp(const cvo_params *p, size_t n1, const double *pos1, size_t n2,
                        const double *pos2, double *grad1, double *grad2, unsigned char *pairlist,
                        int flags, int nthreads)
{
  long double value = 0.0L;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel reduction(+ : value)
  {
    double *g2loc = (flags & CVO_EF_GRADIENTS) ? (double *)calloc(3 * n2, sizeof(double)) : NULL;
#pragma omp for schedule(static)
    for (long long ii = 0; ii < (long long)n1; ii++) {
      size_t const i = (size_t)ii;
      double const a1[3] = {pos1[i], pos1[n1 + i], pos1[2 * n1 + i]};
      double g1[3] = {0, 0, 0};
      for (size_t j = 0; j < n2; j++) {
        double const a2[3] = {pos2[j], pos2[n2 + j], pos2[2 * n2 + j]};
        double g2[3] = {0, 0, 0};
        unsigned char *pl = pairlist ? pairlist + i * n2 + j : NULL;
        int const within = ((flags & CVO_EF_USE_PAIRLIST) &&
                            (*pl || (flags & CVO_EF_REBUILD_PAIRLIST))) ||
                           !(flags & CVO_EF_USE_PAIRLIST);
        double const partial =
            within ? cvo_compute_pair_coordnum(p, a1, a2, g1, g2, flags) : 0.0;
        if ((flags & CVO_EF_USE_PAIRLIST) && (flags & CVO_EF_REBUILD_PAIRLIST)) {
          *pl = partial > 0.0 ? 1 : 0;
        }
        value += partial;
        if (g2loc) {
          g2loc[j] += g2[0];
          g2loc[n2 + j] += g2[1];
          g2loc[2 * n2 + j] += g2[2];
        }
      }
      if (flags & CVO_EF_GRADIENTS) {
        grad1[i] += g1[0];
        grad1[n1 + i] += g1[1];
        grad1[2 * n1 + i] += g1[2];
      }
    }
    if (g2loc) {
#pragma omp critical
      for (size_t k = 0; k < 3 * n2; k++) grad2[k] += g2loc[k];
      free(g2loc);
    }
  }
  return (double)value;
}

double cvo_selfcoordnum_omp(const cvo_params *p, size_t n, const double *pos, double *grad,
                            unsigned char *pairlist, int flags, int nthreads)
{
  long double value = 0.0L;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel reduction(+ : value)
  {
    double *gloc = (flags & CVO_EF_GRADIENTS) ? (double *)calloc(3 * n, sizeof(double)) : NULL;
#pragma omp for schedule(dynamic, 16)
    for (long long ii = 0; ii < (long long)n - 1; ii++) {
      size_t const i = (size_t)ii;
      double const a1[3] = {pos[i], pos[n + i], pos[2 * n + i]};
      double g1[3] = {0, 0, 0};
      /* canonical row-major upper-triangle offset of (i, i+1) */
      size_t const row0 = i * (2 * n - i - 1) / 2;
      for (size_t j = i + 1; j < n; j++) {
        double const a2[3] = {pos[j], pos[n + j], pos[2 * n + j]};
        double g2[3] = {0, 0, 0};
        unsigned char *pl = pairlist ? pairlist + row0 + (j - i - 1) : NULL;
        int const within = ((flags & CVO_EF_USE_PAIRLIST) &&
                            (*pl || (flags & CVO_EF_REBUILD_PAIRLIST))) ||
                           !(flags & CVO_EF_USE_PAIRLIST);
        double const partial =
            within ? cvo_compute_pair_coordnum(p, a1, a2, g1, g2, flags) : 0.0;
        if ((flags & CVO_EF_USE_PAIRLIST) && (flags & CVO_EF_REBUILD_PAIRLIST)) {
          *pl = partial > 0.0 ? 1 : 0;
        }
        value += partial;
        if (gloc) {
          gloc[j] += g2[0];
          gloc[n + j] += g2[1];
          gloc[2 * n + j] += g2[2];
        }
      }
      if (gloc) {
        gloc[i] += g1[0];
        gloc[n + i] += g1[1];
        gloc[2 * n + i] += g1[2];
      }
    }
    if (gloc) {
#pragma omp critical
      for (size_t k = 0; k < 3 * n; k++) grad[k] += gloc[k];
      free(gloc);
    }
  }
  return (double)value;
}

/* ---- atom groups ---------------------------------------------------------------- */

/* src/colvaratoms.cpp:1107-1125; CPU proxy positions are AoS rvector (src/colvarproxy.h:275) */
void cvo_read_positions(size_t n, const int *atoms_index, const double *proxy_pos_aos,
                        double *pos_soa)
{
  for (size_t i = 0; i < n; i++) {
    int const pi = atoms_index[i];
    pos_soa[i] = proxy_pos_aos[3 * pi + 0];
    pos_soa[n + i] = proxy_pos_aos[3 * pi + 1];
    pos_soa[2 * n + i] = proxy_pos_aos[3 * pi + 2];
  }
}

/* src/colvaratoms.cpp:1354-1372 */
void cvo_center_of_geometry(size_t n, const double *pos_soa, double cog[3])
{
  cog[0] = cog[1] = cog[2] = 0.0;
  for (size_t i = 0; i < n; i++) {
    cog[0] += pos_soa[i];
    cog[1] += pos_soa[n + i];
    cog[2] += pos_soa[2 * n + i];
  }
  cog[0] /= (double)n;
  cog[1] /= (double)n;
  cog[2] /= (double)n;
}

/* src/colvaratoms.cpp:1374-1395 */
void cvo_center_of_mass(size_t n, const double *pos_soa, const double *mass, double total_mass,
                        double com[3])
{
  com[0] = com[1] = com[2] = 0.0;
  for (size_t i = 0; i < n; i++) {
    com[0] += mass[i] * pos_soa[i];
    com[1] += mass[i] * pos_soa[n + i];
    com[2] += mass[i] * pos_soa[2 * n + i];
  }
  com[0] /= total_mass;
  com[1] /= total_mass;
  com[2] /= total_mass;
}

/* src/colvaratoms.cpp:1637-1689; proxy->apply_atom_force accumulates (src/colvarproxy.h:139-142) */
void cvo_apply_colvar_force(size_t n, const int *atoms_index, const double *grad_soa,
                            double force, const double *rot_inv, double *proxy_force_aos)
{
  for (size_t i = 0; i < n; i++) {
    int const pi = atoms_index[i];
    double const gx = grad_soa[i], gy = grad_soa[n + i], gz = grad_soa[2 * n + i];
    double fx, fy, fz;
    if (rot_inv) {
      fx = force * (rot_inv[0] * gx + rot_inv[1] * gy + rot_inv[2] * gz);
      fy = force * (rot_inv[3] * gx + rot_inv[4] * gy + rot_inv[5] * gz);
      fz = force * (rot_inv[6] * gx + rot_inv[7] * gy + rot_inv[8] * gz);
    } else {
      fx = force * gx;
      fy = force * gy;
      fz = force * gz;
    }
    proxy_force_aos[3 * pi + 0] += fx;
    proxy_force_aos[3 * pi + 1] += fy;
    proxy_force_aos[3 * pi + 2] += fz;
  }
}


/* ---- distanceInv (widening) ------------------------------- */

/* src/colvarcomp_distances.cpp:405-456 */
double cvo_distance_inv(const cvo_boundaries *bc, int exponent, size_t n1, const double *pos1,
                        size_t n2, const double *pos2, double *grad1, double *grad2,
                        int nthreads)
{
  int const factor = -1 * (exponent / 2);
  double sum = 0.0;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
  if (nthreads <= 1) {
    for (size_t i = 0; i < n1; ++i) {
      double const p1[3] = {pos1[i], pos1[n1 + i], pos1[2 * n1 + i]};
      double g1[3] = {0, 0, 0};
      for (size_t j = 0; j < n2; ++j) {
        double const p2[3] = {pos2[j], pos2[n2 + j], pos2[2 * n2 + j]};
        double dv[3];
        cvo_position_distance(bc, p1, p2, dv);
        double const d2 = dv[0] * dv[0] + dv[1] * dv[1] + dv[2] * dv[2];
        double const dinv = cvo_integer_power(d2, factor);
        sum += dinv;
        double const c = factor * dinv / d2 * 2.0;
        double const gx = c * dv[0], gy = c * dv[1], gz = c * dv[2];
        g1[0] += -1.0 * gx;
        g1[1] += -1.0 * gy;
        g1[2] += -1.0 * gz;
        grad2[j] += gx;
        grad2[n2 + j] += gy;
        grad2[2 * n2 + j] += gz;
      }
      grad1[i] += g1[0];
      grad1[n1 + i] += g1[1];
      grad1[2 * n1 + i] += g1[2];
    }
  } else {
    /* checker for large sizes, NOT the reference's summation order: rows in parallel and the
     * scalar sum Kahan-compensated.  The reference adds ~N1*N2 positive terms of wildly
     * different magnitude one after the other, which loses up to N1*N2*2^-53 relative (terms
     * below half an ulp of the running sum vanish); this variant is what that loop would give
     * in exact arithmetic, to ~1e-16. */
    long double total = 0.0L;
#pragma omp parallel
    {
      double *g2loc = (double *)calloc(3 * n2, sizeof(double));
      long double tsum = 0.0L;
#pragma omp for schedule(static)
      for (long long ii = 0; ii < (long long)n1; ii++) {
        size_t const i = (size_t)ii;
        double const p1[3] = {pos1[i], pos1[n1 + i], pos1[2 * n1 + i]};
        double g1[3] = {0, 0, 0};
        double rs = 0.0, rc = 0.0; /* Kahan pair of this row */
        for (size_t j = 0; j < n2; ++j) {
          double const p2[3] = {pos2[j], pos2[n2 + j], pos2[2 * n2 + j]};
          double dv[3];
          cvo_position_distance(bc, p1, p2, dv);
          double const d2 = dv[0] * dv[0] + dv[1] * dv[1] + dv[2] * dv[2];
          double const dinv = cvo_integer_power(d2, factor);
          double const y = dinv - rc;
          double const t = rs + y;
          rc = (t - rs) - y;
          rs = t;
          double const c = factor * dinv / d2 * 2.0;
          g1[0] += -1.0 * (c * dv[0]);
          g1[1] += -1.0 * (c * dv[1]);
          g1[2] += -1.0 * (c * dv[2]);
          g2loc[j] += c * dv[0];
          g2loc[n2 + j] += c * dv[1];
          g2loc[2 * n2 + j] += c * dv[2];
        }
        tsum += (long double)rs - (long double)rc;
        grad1[i] += g1[0];
        grad1[n1 + i] += g1[1];
        grad1[2 * n1 + i] += g1[2];
      }
#pragma omp critical
      {
        total += tsum;
        for (size_t k = 0; k < 3 * n2; k++) grad2[k] += g2loc[k];
      }
      free(g2loc);
    }
    sum = (double)total;
  }
  double value = sum * (1.0 / (double)(n1 * n2));
  value = pow(value, -1.0 / (double)exponent);
  double const dxdsum =
      (-1.0 / (double)exponent) * cvo_integer_power(value, exponent + 1) / (double)(n1 * n2);
  for (size_t k = 0; k < 3 * n1; k++) grad1[k] = grad1[k] * dxdsum;
  for (size_t k = 0; k < 3 * n2; k++) grad2[k] = grad2[k] * dxdsum;
  return value;
}

/* src/colvarcomp_distances.cpp:486-548 (colvar::distance_pairs): dist[i1*n2 + i2] =
 * |position_distance(p1, p2)|; with force != NULL also the atom forces of apply_force,
 * f2 = force_ij * dv.unit() (src/colvartypes.h:837-841), f1 -= f2, accumulated in loop order. */
void cvo_distance_pairs(const cvo_boundaries *bc, size_t n1, const double *pos1, size_t n2,
                        const double *pos2, double *dist, const double *force, double *f1,
                        double *f2)
{
  for (size_t i1 = 0; i1 < n1; i1++) {
    double const a1[3] = {pos1[i1], pos1[n1 + i1], pos1[2 * n1 + i1]};
    double g1[3] = {0, 0, 0};
    for (size_t i2 = 0; i2 < n2; i2++) {
      double const a2[3] = {pos2[i2], pos2[n2 + i2], pos2[2 * n2 + i2]};
      double dv[3];
      cvo_position_distance(bc, a1, a2, dv);
      double const d = sqrt(dv[0] * dv[0] + dv[1] * dv[1] + dv[2] * dv[2]);
      dist[i1 * n2 + i2] = d;
      if (force) {
        double u[3] = {1.0, 0.0, 0.0};
        if (d > 0.0) {
          u[0] = dv[0] / d;
          u[1] = dv[1] / d;
          u[2] = dv[2] / d;
        }
        for (int k = 0; k < 3; k++) {
          double const f = force[i1 * n2 + i2] * u[k];
          g1[k] += -f;
          f2[k * n2 + i2] += f;
        }
      }
    }
    if (force) {
      for (int k = 0; k < 3; k++) f1[k * n1 + i1] += g1[k];
    }
  }
}

