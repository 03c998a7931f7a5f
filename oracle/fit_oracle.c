/*
 * fit_oracle.c -- TEST INFRASTRUCTURE ONLY (see coordnum_oracle.h).
 *
 * CPU restatement of the fitting path of an atom group used by coordNum config 4:
 *   optimal rotation     src/colvartypes.cpp:211-255 (correlation / overlap matrix),
 *                        :398-489 (calc_optimal_rotation_soa / _impl), eigen-solve of the
 *                        4x4 overlap matrix (src/nr_jacobi.cpp; any convergent Jacobi gives
 *                        the same eigenpairs up to sign -- this one is a plain full-matrix
 *                        cyclic Jacobi)
 *   roto-translation     src/colvaratoms.cpp:1127-1173 (calc_apply_roto_translation)
 *   fit gradients        src/colvaratoms.cpp:1471-1527 (calc_fit_forces_impl) with the
 *                        quaternion derivative of src/colvar_rotation_derivative.h:97-330
 *                        (first-order eigenvector perturbation) and :535-570 (dS/dC chain),
 *                        dR/dq from src/colvartypes.h:1249-1264.
 * Pinned against outputs of the unmodified reference (tests/golden/ref_vectors/fit_*.npz).
 */
#include "coordnum_oracle.h"

#include <math.h>
#include <string.h>

/* symmetric 4x4 eigen-decomposition, cyclic Jacobi; eigenvalues descending, eigenvectors as
 * rows of `vec`, normalised */
static void jacobi4(const double Sin[4][4], double val[4], double vec[4][4])
{
  double A[4][4], V[4][4];
  memcpy(A, Sin, sizeof(A));
  memset(V, 0, sizeof(V));
  for (int i = 0; i < 4; i++) V[i][i] = 1.0;
  for (int sweep = 0; sweep < 64; sweep++) {
    double off = 0.0, diag = 0.0;
    for (int p = 0; p < 4; p++) {
      diag += A[p][p] * A[p][p];
      for (int q = p + 1; q < 4; q++) off += A[p][q] * A[p][q];
    }
    if (off <= 1e-34 * diag || off == 0.0) break;
    for (int p = 0; p < 3; p++) {
      for (int q = p + 1; q < 4; q++) {
        double const apq = A[p][q];
        if (apq == 0.0) continue;
        double const theta = (A[q][q] - A[p][p]) / (2.0 * apq);
        double t = 1.0 / (fabs(theta) + sqrt(theta * theta + 1.0));
        if (theta < 0.0) t = -t;
        double const c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        /* A <- J^T A J with J the rotation in the (p,q) plane */
        for (int k = 0; k < 4; k++) {
          double const akp = A[k][p], akq = A[k][q];
          A[k][p] = c * akp - s * akq;
          A[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < 4; k++) {
          double const apk = A[p][k], aqk = A[q][k];
          A[p][k] = c * apk - s * aqk;
          A[q][k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 4; k++) {
          double const vkp = V[k][p], vkq = V[k][q];
          V[k][p] = c * vkp - s * vkq;
          V[k][q] = s * vkp + c * vkq;
        }
      }
    }
  }
  int order[4] = {0, 1, 2, 3};
  for (int i = 0; i < 4; i++) {
    for (int j = i + 1; j < 4; j++) {
      if (A[order[j]][order[j]] > A[order[i]][order[i]]) {
        int const t = order[i];
        order[i] = order[j];
        order[j] = t;
      }
    }
  }
  for (int e = 0; e < 4; e++) {
    int const col = order[e];
    val[e] = A[col][col];
    double n2 = 0.0;
    for (int k = 0; k < 4; k++) n2 += V[k][col] * V[k][col];
    double const inv = 1.0 / sqrt(n2);
    for (int k = 0; k < 4; k++) vec[e][k] = V[k][col] * inv;
  }
}

/* src/colvartypes.cpp:231-255: overlap matrix from the correlation matrix */
static void overlap_from_correlation(const double C[3][3], double S[4][4])
{
  S[0][0] = C[0][0] + C[1][1] + C[2][2];
  S[1][0] = S[0][1] = C[1][2] - C[2][1];
  S[2][0] = S[0][2] = -C[0][2] + C[2][0];
  S[3][0] = S[0][3] = C[0][1] - C[1][0];
  S[1][1] = C[0][0] - C[1][1] - C[2][2];
  S[2][1] = S[1][2] = C[0][1] + C[1][0];
  S[3][1] = S[1][3] = C[0][2] + C[2][0];
  S[2][2] = -C[0][0] + C[1][1] - C[2][2];
  S[3][2] = S[2][3] = C[1][2] + C[2][1];
  S[3][3] = -C[0][0] - C[1][1] + C[2][2];
}

void cvo_optimal_rotation(size_t n, const double *pos, const double *ref, double q[4],
                          double eigval[4], double eigvec[4][4], double S[4][4])
{
  double C[3][3] = {{0}};
  for (size_t i = 0; i < n; i++) {
    for (int a = 0; a < 3; a++) {
      for (int b = 0; b < 3; b++) C[a][b] += pos[a * n + i] * ref[b * n + i];
    }
  }
  overlap_from_correlation(C, S);
  jacobi4(S, eigval, eigvec);
  memcpy(q, eigvec[0], 4 * sizeof(double));
}

/* cvm::quaternion::rotation_matrix, src/colvartypes.h:1194-1213 */
void cvo_quaternion_to_matrix(const double q[4], double R[3][3])
{
  double const q0 = q[0], q1 = q[1], q2 = q[2], q3 = q[3];
  R[0][0] = q0 * q0 + q1 * q1 - q2 * q2 - q3 * q3;
  R[1][1] = q0 * q0 - q1 * q1 + q2 * q2 - q3 * q3;
  R[2][2] = q0 * q0 - q1 * q1 - q2 * q2 + q3 * q3;
  R[0][1] = 2.0 * (q1 * q2 - q0 * q3);
  R[0][2] = 2.0 * (q0 * q2 + q1 * q3);
  R[1][0] = 2.0 * (q0 * q3 + q1 * q2);
  R[1][2] = 2.0 * (q2 * q3 - q0 * q1);
  R[2][0] = 2.0 * (q1 * q3 - q0 * q2);
  R[2][1] = 2.0 * (q0 * q1 + q2 * q3);
}

static void translate(size_t n, double *pos, const double t[3], double s)
{
  for (size_t i = 0; i < n; i++) {
    pos[i] += s * t[0];
    pos[n + i] += s * t[1];
    pos[2 * n + i] += s * t[2];
  }
}

static void rotate(size_t n, double *pos, const double R[3][3])
{
  for (size_t i = 0; i < n; i++) {
    double const x = pos[i], y = pos[n + i], z = pos[2 * n + i];
    pos[i] = R[0][0] * x + R[0][1] * y + R[0][2] * z;
    pos[n + i] = R[1][0] * x + R[1][1] * y + R[1][2] * z;
    pos[2 * n + i] = R[2][0] * x + R[2][1] * y + R[2][2] * z;
  }
}

/* src/colvaratoms.cpp:1127-1173.  pos_fit == NULL: the main group is fitted on itself. */
void cvo_apply_roto_translation(cvo_fit_state *fs, size_t nmain, double *pos_main,
                                double *pos_main_unrotated, size_t nfit, double *pos_fit,
                                const double *ref_pos_centered)
{
  double *fit = pos_fit ? pos_fit : pos_main;
  size_t const nf = pos_fit ? nfit : nmain;
  if (fs->center) {
    double cog[3];
    cvo_center_of_geometry(nf, fit, cog);
    translate(nmain, pos_main, cog, -1.0);
    if (pos_fit) translate(nfit, pos_fit, cog, -1.0);
  }
  if (pos_main_unrotated) memcpy(pos_main_unrotated, pos_main, 3 * nmain * sizeof(double));
  if (fs->rotate) {
    cvo_optimal_rotation(nf, fit, ref_pos_centered, fs->q, fs->eigval, fs->eigvec, fs->S);
    cvo_quaternion_to_matrix(fs->q, fs->R);
    rotate(nmain, pos_main, fs->R);
    if (pos_fit) rotate(nfit, pos_fit, fs->R);
  }
  if (fs->center && !fs->center_origin) {
    translate(nmain, pos_main, fs->ref_pos_cog, 1.0);
    if (pos_fit) translate(nfit, pos_fit, fs->ref_pos_cog, 1.0);
  }
}

/* src/colvaratoms.cpp:1471-1527 */
void cvo_calc_fit_gradients(const cvo_fit_state *fs, size_t nmain, const double *grad_main,
                            const double *pos_main_unrotated, size_t nfit,
                            const double *pos_fit_centered, const double *ref_pos_centered,
                            double *fit_grad)
{
  (void)pos_fit_centered;
  double sum_g[3] = {0, 0, 0};
  double Cg[3][3] = {{0}};
  for (size_t i = 0; i < nmain; i++) {
    double const g[3] = {grad_main[i], grad_main[nmain + i], grad_main[2 * nmain + i]};
    if (fs->center) {
      sum_g[0] += g[0];
      sum_g[1] += g[1];
      sum_g[2] += g[2];
    }
    if (fs->rotate) {
      for (int a = 0; a < 3; a++) {
        for (int b = 0; b < 3; b++) Cg[a][b] += g[a] * pos_main_unrotated[b * nmain + i];
      }
    }
  }
  double atom_grad[3] = {0, 0, 0};
  if (fs->center) {
    if (fs->rotate) {
      /* rot.inverse() = transpose */
      for (int a = 0; a < 3; a++) {
        atom_grad[a] = fs->R[0][a] * sum_g[0] + fs->R[1][a] * sum_g[1] + fs->R[2][a] * sum_g[2];
      }
    } else {
      memcpy(atom_grad, sum_g, sizeof(sum_g));
    }
    for (int a = 0; a < 3; a++) atom_grad[a] *= (-1.0) / (double)nfit;
  }
  double dxdC[3][3] = {{0}};
  if (fs->rotate) {
    double const q0 = fs->q[0], q1 = fs->q[1], q2 = fs->q[2], q3 = fs->q[3];
    /* dR/dq_p, src/colvartypes.h:1249-1264 */
    double const dR[4][3][3] = {
        {{q0, -q3, q2}, {q3, q0, -q1}, {-q2, q1, q0}},
        {{q1, q2, q3}, {q2, -q1, -q0}, {q3, q0, -q1}},
        {{-q2, q1, q0}, {q1, q2, q3}, {-q0, q3, -q2}},
        {{-q3, -q0, q1}, {q0, -q3, q2}, {q1, q2, q3}}};
    double dxdq[4];
    for (int p = 0; p < 4; p++) {
      double s = 0.0;
      for (int a = 0; a < 3; a++) {
        for (int b = 0; b < 3; b++) s += dR[p][a][b] * Cg[a][b];
      }
      dxdq[p] = 2.0 * s;
    }
    /* T[p][i][j] = dq_p / dS_ij = sum_k Qk[p] Qk[i] Q0[j] / (L0 - Lk)
     * (src/colvar_rotation_derivative.h:125-329) */
    double T[4][4][4];
    for (int p = 0; p < 4; p++) {
      for (int i = 0; i < 4; i++) {
        for (int j = 0; j < 4; j++) {
          double s = 0.0;
          for (int k = 1; k < 4; k++) {
            s += (fs->eigvec[k][i] * fs->eigvec[0][j]) / (fs->eigval[0] - fs->eigval[k]) *
                 fs->eigvec[k][p];
          }
          T[p][i][j] = s;
        }
      }
    }
    /* chain through dS/dC (src/colvar_rotation_derivative.h:535-548) */
    for (int p = 0; p < 4; p++) {
      double const f = dxdq[p];
      dxdC[0][0] += f * (T[p][0][0] + T[p][1][1] - T[p][2][2] - T[p][3][3]);
      dxdC[0][1] += f * (T[p][0][3] + T[p][1][2] + T[p][2][1] + T[p][3][0]);
      dxdC[0][2] += f * (-T[p][0][2] + T[p][1][3] - T[p][2][0] + T[p][3][1]);
      dxdC[1][0] += f * (-T[p][0][3] + T[p][1][2] + T[p][2][1] - T[p][3][0]);
      dxdC[1][1] += f * (T[p][0][0] - T[p][1][1] + T[p][2][2] - T[p][3][3]);
      dxdC[1][2] += f * (T[p][0][1] + T[p][1][0] + T[p][2][3] + T[p][3][2]);
      dxdC[2][0] += f * (T[p][0][2] + T[p][1][3] + T[p][2][0] + T[p][3][1]);
      dxdC[2][1] += f * (-T[p][0][1] - T[p][1][0] + T[p][2][3] + T[p][3][2]);
      dxdC[2][2] += f * (T[p][0][0] - T[p][1][1] - T[p][2][2] + T[p][3][3]);
    }
  }
  for (size_t j = 0; j < nfit; j++) {
    double g[3] = {0, 0, 0};
    if (fs->center) {
      g[0] += atom_grad[0];
      g[1] += atom_grad[1];
      g[2] += atom_grad[2];
    }
    if (fs->rotate) {
      double const rx = ref_pos_centered[j], ry = ref_pos_centered[nfit + j],
                   rz = ref_pos_centered[2 * nfit + j];
      g[0] += dxdC[0][0] * rx + dxdC[0][1] * ry + dxdC[0][2] * rz;
      g[1] += dxdC[1][0] * rx + dxdC[1][1] * ry + dxdC[1][2] * rz;
      g[2] += dxdC[2][0] * rx + dxdC[2][1] * ry + dxdC[2][2] * rz;
    }
    fit_grad[j] = g[0];
    fit_grad[nfit + j] = g[1];
    fit_grad[2 * nfit + j] = g[2];
  }
}
