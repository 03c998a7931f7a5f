// tiles.cu -- see tiles.cuh.  Grid and sort of both groups, cluster table, the tile build
// (rebuild step) and the tile walk (list steps).
#include "tiles.cuh"

#include <algorithm>
#include <cmath>

#include <cub/device/device_radix_sort.cuh>

#include "kernels.cuh"

namespace b200cv {

namespace {

constexpr int GRID_THREADS = 1024;
constexpr int WARPS = TILE_THREADS / 32;

// ---- bounding box of group2 -> grid with cells about as large as the volume of 32 atoms ------
__global__ void __launch_bounds__(GRID_THREADS)
tile_grid_kernel(const double *__restrict__ x, const double *__restrict__ y,
                 const double *__restrict__ z, int n, TileGrid *grid, const int *gate)
{
  __shared__ double red[6][GRID_THREADS / 32];
  if (gate != nullptr && *gate != 1) return;
  double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
  for (int i = threadIdx.x; i < n; i += GRID_THREADS) {
    double const p[3] = {x[i], y[i], z[i]};
#pragma unroll
    for (int d = 0; d < 3; d++) {
      lo[d] = fmin(lo[d], p[d]);
      hi[d] = fmax(hi[d], p[d]);
    }
  }
#pragma unroll
  for (int d = 0; d < 3; d++) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
      lo[d] = fmin(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], s));
      hi[d] = fmax(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], s));
    }
    if ((threadIdx.x & 31) == 0) {
      red[d][threadIdx.x >> 5] = lo[d];
      red[3 + d][threadIdx.x >> 5] = hi[d];
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a[3], ext[3], emax = 0.0;
    for (int d = 0; d < 3; d++) {
      double l = red[d][0], h = red[3 + d][0];
      for (int w = 1; w < GRID_THREADS / 32; w++) {
        l = fmin(l, red[d][w]);
        h = fmax(h, red[3 + d][w]);
      }
      a[d] = l;
      ext[d] = fmax(h - l, 0.0);
      emax = fmax(emax, ext[d]);
    }
    // cell edge: the cube that holds 32 atoms at the mean density of the occupied box
    // (flat or linear systems: only the axes that have an extent count)
    int k = 0;
    double vol = 1.0;
    for (int d = 0; d < 3; d++) {
      if (ext[d] > 1.0e-9 * emax && ext[d] > 0.0) {
        vol *= ext[d];
        k++;
      }
    }
    double s = 1.0;
    if (k > 0) s = pow(32.0 * vol / (double)max(n, 1), 1.0 / (double)k);
    int ncells = 1;
    for (int d = 0; d < 3; d++) {
      double cell = fmax(s, ext[d] / (double)TILE_MAX_DIM * (1.0 + 1.0e-9));
      if (!(cell > 0.0)) cell = 1.0;
      int dim = (int)floor(ext[d] / cell) + 1;
      dim = max(1, min(dim, TILE_MAX_DIM));
      grid->origin[d] = a[d];
      grid->inv_cell[d] = 1.0 / cell;
      grid->dim[d] = dim;
      ncells *= dim;
    }
    grid->ncells = ncells;
    grid->nrows = grid->dim[1] * grid->dim[2];
  }
}

__global__ void tile_key_kernel(const double *__restrict__ x, const double *__restrict__ y,
                                const double *__restrict__ z, int n,
                                const TileGrid *__restrict__ grid, int *key, int *val,
                                const int *gate)
{
  int const i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (gate != nullptr && *gate != 1) return;  // keys of the last rebuild stay: still sortable
  TileGrid const g = *grid;
  int const cx = tile_cell(x[i], g.origin[0], g.inv_cell[0], g.dim[0]);
  int const cy = tile_cell(y[i], g.origin[1], g.inv_cell[1], g.dim[1]);
  int const cz = tile_cell(z[i], g.origin[2], g.inv_cell[2], g.dim[2]);
  // position inside the cell along x, TILE_FINE_BITS bits: consecutive atoms of a row of cells
  // are neighbours along x as well
  double const fx = (x[i] - g.origin[0]) * g.inv_cell[0] - (double)cx;
  int const fine = (int)fmin(fmax(fx * (double)(1 << TILE_FINE_BITS), 0.0),
                             (double)((1 << TILE_FINE_BITS) - 1));
  key[i] = ((((cz * g.dim[1] + cy) * g.dim[0]) + cx) << TILE_FINE_BITS) | fine;
  val[i] = i;
}

// cell_start[c] = first sorted position whose cell is >= c, for c in [0, ncells]
__global__ void tile_cell_start_kernel(const int *__restrict__ key_sorted, int n,
                                       const TileGrid *__restrict__ grid, int *cell_start,
                                       const int *gate)
{
  int const c = blockIdx.x * blockDim.x + threadIdx.x;
  if (gate != nullptr && *gate != 1) return;
  if (c > grid->ncells) return;
  int lo = 0, hi = n;
  while (lo < hi) {
    int const mid = (lo + hi) >> 1;
    if ((key_sorted[mid] >> TILE_FINE_BITS) < c) lo = mid + 1;
    else hi = mid;
  }
  cell_start[c] = lo;
}

// clusters of 32 consecutive sorted atoms, restarted at every row of cells (one CTA)
__global__ void __launch_bounds__(1024)
tile_cluster_kernel(const int *__restrict__ cell_start, const TileGrid *__restrict__ grid,
                    int4 *clusters, int *ncl, int max_clusters, const int *gate)
{
  __shared__ int scan[2][1024];
  if (gate != nullptr && *gate != 1) return;
  int const nrows = grid->nrows, dim0 = grid->dim[0];
  int const per = (nrows + 1023) / 1024;
  int const r0 = min((int)threadIdx.x * per, nrows), r1 = min(r0 + per, nrows);
  int mine = 0;
  for (int r = r0; r < r1; r++) {
    int const cnt = cell_start[(r + 1) * dim0] - cell_start[r * dim0];
    mine += (cnt + 31) >> 5;
  }
  scan[0][threadIdx.x] = mine;
  __syncthreads();
  int cur = 0;
  for (int d = 1; d < 1024; d <<= 1) {
    int v = scan[cur][threadIdx.x];
    if ((int)threadIdx.x >= d) v += scan[cur][threadIdx.x - d];
    scan[cur ^ 1][threadIdx.x] = v;
    cur ^= 1;
    __syncthreads();
  }
  int base = scan[cur][threadIdx.x] - mine;  // exclusive
  if (threadIdx.x == 1023) *ncl = min(scan[cur][1023], max_clusters);
  for (int r = r0; r < r1; r++) {
    int const p = cell_start[r * dim0];
    int const cnt = cell_start[(r + 1) * dim0] - p;
    for (int k = 0; k < cnt; k += 32) {
      if (base < max_clusters) clusters[base] = make_int4(p + k, min(32, cnt - k), r, 0);
      base++;
    }
  }
}

// coordinates into sorted order + sorted gradient planes zeroed; both sides in one launch
__global__ void tile_gather_kernel(const double *__restrict__ x1, const double *__restrict__ y1,
                                   const double *__restrict__ z1, const int *__restrict__ s1,
                                   int n1, double *xs1, double *ys1, double *zs1, double *gs1,
                                   const double *__restrict__ x2, const double *__restrict__ y2,
                                   const double *__restrict__ z2, const int *__restrict__ s2,
                                   int n2, double *xs2, double *ys2, double *zs2, double *gs2,
                                   int blocks1, const int *gate, int gate_on)
{
  if (gate != nullptr && *gate != gate_on) return;
  if ((int)blockIdx.x < blocks1) {
    int const p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n1) return;
    int const i = s1[p];
    xs1[p] = x1[i];
    ys1[p] = y1[i];
    zs1[p] = z1[i];
    gs1[p] = 0.0;
    gs1[n1 + p] = 0.0;
    gs1[2 * (size_t)n1 + p] = 0.0;
  } else {
    int const p = ((int)blockIdx.x - blocks1) * blockDim.x + threadIdx.x;
    if (p >= n2) return;
    int const j = s2[p];
    xs2[p] = x2[j];
    ys2[p] = y2[j];
    zs2[p] = z2[j];
    gs2[p] = 0.0;
    gs2[n2 + p] = 0.0;
    gs2[2 * (size_t)n2 + p] = 0.0;
  }
}

// every atom appears once in a sorted order: plain read-modify-write
__global__ void tile_scatter_add_kernel(const double *__restrict__ gs1, const int *__restrict__ s1,
                                        int n1, double *g1, size_t st1,
                                        const double *__restrict__ gs2, const int *__restrict__ s2,
                                        int n2, double *g2, size_t st2, int blocks1,
                                        const int *gate, int gate_on)
{
  if (gate != nullptr && *gate != gate_on) return;
  if ((int)blockIdx.x < blocks1) {
    int const p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n1) return;
    int const i = s1[p];
    g1[i] += gs1[p];
    g1[st1 + i] += gs1[n1 + p];
    g1[2 * st1 + i] += gs1[2 * (size_t)n1 + p];
  } else {
    int const p = ((int)blockIdx.x - blocks1) * blockDim.x + threadIdx.x;
    if (p >= n2) return;
    int const j = s2[p];
    g2[j] += gs2[p];
    g2[st2 + j] += gs2[n2 + p];
    g2[2 * st2 + j] += gs2[2 * (size_t)n2 + p];
  }
}

// ---- one 32 x 32 tile ---------------------------------------------------------------------------
// BUILD: `mask` comes in as the valid-pair bits (i-atom valid, slot valid, and for selfCoordNum
// the slot's atom behind the i-atom in sorted order) and goes out as the listed-pair bits.
// walk: `mask` holds the listed-pair bits.  Lane l meets slot (l + t) & 31 at step t; b2* travel
// with the slot, so after the loop lane l holds the warp's sum for slot l.
template <int EXP, bool ANISO, bool BUILD>
__device__ __forceinline__ void tile_body(PairParams const &P, int lane, double xi, double yi,
                                          double zi, const double *__restrict__ sx,
                                          const double *__restrict__ sy,
                                          const double *__restrict__ sz, unsigned &mask,
                                          double &a1x, double &a1y, double &a1z, double &fsum,
                                          double &b2x, double &b2y, double &b2z, int pi,
                                          const int *__restrict__ sidx, unsigned long long *rare,
                                          unsigned rare_cap, unsigned *rare_count)
{
  unsigned const in = mask;
  unsigned out = 0u;
  b2x = b2y = b2z = 0.0;
  constexpr int UNROLL = B200CV_TILE_UNROLL;
#pragma unroll UNROLL
  for (int t = 0; t < 32; t++) {
    int const jj = (lane + t) & 31;
    double dx = sx[jj] - xi;
    double dy = sy[jj] - yi;
    double dz = sz[jj] - zi;
    double const u = pair_l2<ANISO, BC_NONE>(P, dx, dy, dz);
    // everything below is branch-free on purpose: the mask bit differs from lane to lane, and a
    // branch on it (which the compiler otherwise builds around the evaluation of u) serialises
    // the steps of a warp.  on = all ones / all zeros from the mask bit.
    int const uh = __double2hiint(u);
    int const on = -(int)((in >> t) & 1u);
    bool const near_thr = (((unsigned)(uh - 0x3FEE0000) < 0x30000u) ||
                           ((unsigned)(uh - P.thr_lo_hi) <= P.thr_width)) && (on != 0);
    int const keep = near_thr ? 0 : on;
    // pair switched off: l2 = 2^100 (B200CV_L2_OFF), beyond any tolerance_l2_max
    double const ue = __hiloint2double((uh & keep) | (0x46300000 & ~keep), __double2loint(u) & keep);
    double F, gq;
    switching_fast<EXP, true>(ue, false, P, F, gq);
    if constexpr (BUILD) out |= (__double_as_longlong(F) != 0ll ? 1u : 0u) << t;
    fsum += F;
    a1x = fma(gq, dx, a1x);
    b2x = fma(gq, dx, b2x);
    a1y = fma(gq, dy, a1y);
    b2y = fma(gq, dy, b2y);
    a1z = fma(gq, dz, a1z);
    b2z = fma(gq, dz, b2z);
    if (__builtin_expect(near_thr, 0)) {
      // next to a decision threshold: reference-order arithmetic in exact_kernel
      unsigned const slot = atomicAdd(rare_count, 1u);
      if (slot < rare_cap) {
        rare[slot] = ((unsigned long long)(unsigned)pi << 32) | (unsigned long long)(unsigned)sidx[jj];
      }
    }
    int const src = (lane + 1) & 31;
    b2x = __shfl_sync(0xffffffffu, b2x, src);
    b2y = __shfl_sync(0xffffffffu, b2y, src);
    b2z = __shfl_sync(0xffffffffu, b2z, src);
  }
  if constexpr (BUILD) mask = out;
}

__device__ __forceinline__ double warp_min(double v)
{
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, s));
  return v;
}
__device__ __forceinline__ double warp_max(double v)
{
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, s));
  return v;
}

struct WarpStage {
  int idx[64];
  double x[64], y[64], z[64];
};

// ---- rebuild: one warp per i-cluster ------------------------------------------------------------
template <int EXP, bool ANISO>
__global__ void __launch_bounds__(TILE_THREADS) tile_build_kernel(const __grid_constant__ TileBuildArgs A)
{
  __shared__ WarpStage stage[WARPS];
  __shared__ double red[WARPS];
  __shared__ unsigned long long redn[WARPS];
  if (A.gate != nullptr && *A.gate != 1) {
    if (threadIdx.x == 0) A.partials[blockIdx.x] = 0.0;
    return;
  }
  int const lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  WarpStage &S = stage[warp];
  // stale slots are read (and switched off by their mask bit): they must hold finite numbers
  for (int k = lane; k < 64; k += 32) {
    S.idx[k] = 0;
    S.x[k] = S.y[k] = S.z[k] = 0.0;
  }
  __syncwarp();
  TileGrid const g = *A.grid;
  PairParams const &P = A.P;
  int const ncl = *A.ncl;
  int const c_begin = (int)((long long)ncl * A.rank / A.world);
  int const c_end = (int)((long long)ncl * (A.rank + 1) / A.world);
  double const l2_pass = P.l2_max * (1.0 + 1.0e-9);
  double fsum = 0.0;
  unsigned long long npairs = 0ull;
  unsigned *const unit_count = A.counts;
  unsigned *const ticket = A.counts + 1;

  for (;;) {
    int c = 0;
    if (lane == 0) c = c_begin + (int)atomicAdd(ticket, 1u);
    c = __shfl_sync(0xffffffffu, c, 0);
    if (c >= c_end) break;
    int4 const cl = A.clusters[c];
    int const p0 = cl.x, cnt = cl.y, row_i = cl.z;
    bool const ivalid = lane < cnt;
    int const pi = ivalid ? p0 + lane : p0;
    double const xi = A.xi[pi], yi = A.yi[pi], zi = A.zi[pi];
    double const xmin = warp_min(xi), xmax = warp_max(xi);
    double const ymin = warp_min(yi), ymax = warp_max(yi);
    double const zmin = warp_min(zi), zmax = warp_max(zi);
    int const ix_lo = tile_cell(xmin - A.cut[0], g.origin[0], g.inv_cell[0], g.dim[0]);
    int const ix_hi = tile_cell(xmax + A.cut[0], g.origin[0], g.inv_cell[0], g.dim[0]);
    int const iy_lo = tile_cell(ymin - A.cut[1], g.origin[1], g.inv_cell[1], g.dim[1]);
    int const iy_hi = tile_cell(ymax + A.cut[1], g.origin[1], g.inv_cell[1], g.dim[1]);
    int const iz_lo = tile_cell(zmin - A.cut[2], g.origin[2], g.inv_cell[2], g.dim[2]);
    int const iz_hi = tile_cell(zmax + A.cut[2], g.origin[2], g.inv_cell[2], g.dim[2]);

    double a1x = 0.0, a1y = 0.0, a1z = 0.0;
    int nst = 0;        // staged candidates (warp-uniform)
    int unit = -1;      // unit being filled (warp-uniform)
    int ntiles = 0;

    auto close_unit = [&]() {
      if (unit >= 0 && (unsigned)unit < A.cap_units && lane == 0) {
        A.unit_hdr[unit] = make_int4(p0, cnt, ntiles, 0);
      }
    };
    // evaluate the first nv staged candidates against the cluster and emit the tile
    auto process_tile = [&](int nv) {
      unsigned const jvalid = nv >= 32 ? 0xffffffffu : ((1u << nv) - 1u);
      // bit t: slot (lane + t) & 31 valid
      unsigned mask = ivalid ? __funnelshift_r(jvalid, jvalid, lane) : 0u;
      if (A.self) {
        // a pair belongs to the atom that comes first in sorted order
        unsigned later = 0u;
#pragma unroll 4
        for (int t = 0; t < 32; t++) later |= (S.idx[(lane + t) & 31] > pi ? 1u : 0u) << t;
        mask &= later;
      }
      double b2x, b2y, b2z;
      tile_body<EXP, ANISO, true>(P, lane, xi, yi, zi, S.x, S.y, S.z, mask, a1x, a1y, a1z, fsum,
                                  b2x, b2y, b2z, pi, S.idx, A.rare, A.rare_cap, A.rare_count);
      if (A.want_grad && lane < nv && (b2x != 0.0 || b2y != 0.0 || b2z != 0.0)) {
        int const j = S.idx[lane];
        atomicAdd(&A.gj[j], P.c_grad[0] * b2x);
        atomicAdd(&A.gj[(size_t)A.nj + j], P.c_grad[1] * b2y);
        atomicAdd(&A.gj[2 * (size_t)A.nj + j], P.c_grad[2] * b2z);
      }
      if (__ballot_sync(0xffffffffu, mask != 0u) != 0u) {
        if (unit < 0 || ntiles == TILE_UNIT) {
          close_unit();
          int u = 0;
          if (lane == 0) u = (int)atomicAdd(unit_count, 1u);
          unit = __shfl_sync(0xffffffffu, u, 0);
          ntiles = 0;
        }
        if ((unsigned)unit < A.cap_units) {
          size_t const at = ((size_t)unit * TILE_UNIT + ntiles) * 32 + lane;
          A.tile_j[at] = lane < nv ? S.idx[lane] : -1;
          A.tile_m[at] = mask;
        }
        ntiles++;
        npairs += (unsigned long long)__popc(mask);
      }
    };

    for (int iz = iz_lo; iz <= iz_hi; iz++) {
      for (int iy = iy_lo; iy <= iy_hi; iy++) {
        int const row = iz * g.dim[1] + iy;
        if (A.self && row < row_i) continue;  // everything there comes before this cluster
        int begin = A.cell_start[row * g.dim[0] + ix_lo];
        int const end = A.cell_start[row * g.dim[0] + ix_hi + 1];
        if (A.self) begin = max(begin, p0);
        for (int q0 = begin; q0 < end; q0 += 32) {
          int const q = q0 + lane;
          bool pass = false;
          double x = 0.0, y = 0.0, z = 0.0;
          if (q < end) {
            x = A.xj[q];
            y = A.yj[q];
            z = A.zj[q];
            // distance to the cluster's bounding box, in units of the cut-offs
            double const gx = fmax(fmax(xmin - x, x - xmax), 0.0) * P.inv_r0[0];
            double const gy = fmax(fmax(ymin - y, y - ymax), 0.0) * P.inv_r0[1];
            double const gz = fmax(fmax(zmin - z, z - zmax), 0.0) * P.inv_r0[2];
            pass = fma(gz, gz, fma(gy, gy, gx * gx)) <= l2_pass;
          }
          unsigned const m = __ballot_sync(0xffffffffu, pass);
          if (m == 0u) continue;
          if (pass) {
            int const slot = nst + __popc(m & ((1u << lane) - 1u));
            S.idx[slot] = q;
            S.x[slot] = x;
            S.y[slot] = y;
            S.z[slot] = z;
          }
          nst += __popc(m);
          __syncwarp();
          if (nst >= 32) {
            process_tile(32);
            __syncwarp();
            // the rest moves to the front
            int const rest = nst - 32;
            int mi = 0;
            double mx = 0.0, my = 0.0, mz = 0.0;
            if (lane < rest) {
              mi = S.idx[32 + lane];
              mx = S.x[32 + lane];
              my = S.y[32 + lane];
              mz = S.z[32 + lane];
            }
            __syncwarp();
            if (lane < rest) {
              S.idx[lane] = mi;
              S.x[lane] = mx;
              S.y[lane] = my;
              S.z[lane] = mz;
            }
            nst = rest;
            __syncwarp();
          }
        }
      }
    }
    if (nst > 0) {
      process_tile(nst);
      __syncwarp();
    }
    close_unit();
    if (A.want_grad && ivalid) {
      atomicAdd(&A.gi[pi], -P.c_grad[0] * a1x);
      atomicAdd(&A.gi[(size_t)A.ni + pi], -P.c_grad[1] * a1y);
      atomicAdd(&A.gi[2 * (size_t)A.ni + pi], -P.c_grad[2] * a1z);
    }
  }

#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    fsum += __shfl_xor_sync(0xffffffffu, fsum, d);
    npairs += __shfl_xor_sync(0xffffffffu, npairs, d);
  }
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

// ---- list step: one warp per unit -----------------------------------------------------------------
template <int EXP, bool ANISO>
__global__ void __launch_bounds__(TILE_THREADS, B200CV_TILE_MINB) tile_walk_kernel(const __grid_constant__ TileWalkArgs A)
{
  __shared__ double sx[WARPS][32], sy[WARPS][32], sz[WARPS][32];
  __shared__ int sj[WARPS][32];
  __shared__ double red[WARPS];
  if (A.gate != nullptr && *A.gate != 0) {
    if (threadIdx.x == 0) A.partials[blockIdx.x] = 0.0;
    return;
  }
  int const lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  PairParams const &P = A.P;
  unsigned nunits = A.counts[6];
  if (nunits > A.cap_units) nunits = A.cap_units;
  unsigned *const ticket = A.counts + 2;
  double fsum = 0.0;
  unsigned next = 0u;
  if (lane == 0) next = atomicAdd(ticket, 1u);
  for (;;) {
    unsigned const u = __shfl_sync(0xffffffffu, next, 0);
    if (u >= nunits) break;
    if (lane == 0) next = atomicAdd(ticket, 1u);  // in flight while this unit is worked on
    int4 const hdr = A.unit_hdr[u];
    int const p0 = hdr.x, cnt = hdr.y, nt = hdr.z;
    bool const ivalid = lane < cnt;
    int const pi = ivalid ? p0 + lane : p0;
    double const xi = A.xi[pi], yi = A.yi[pi], zi = A.zi[pi];
    double a1x = 0.0, a1y = 0.0, a1z = 0.0;
    for (int k = 0; k < nt; k++) {
      size_t const at = ((size_t)u * TILE_UNIT + k) * 32 + lane;
      int const j = A.tile_j[at];
      unsigned mask = A.tile_m[at];
      int const jl = j < 0 ? 0 : j;
      sx[warp][lane] = A.xj[jl];
      sy[warp][lane] = A.yj[jl];
      sz[warp][lane] = A.zj[jl];
      sj[warp][lane] = jl;
      __syncwarp();
      double b2x, b2y, b2z;
      tile_body<EXP, ANISO, false>(P, lane, xi, yi, zi, sx[warp], sy[warp], sz[warp], mask, a1x,
                                   a1y, a1z, fsum, b2x, b2y, b2z, pi, sj[warp], A.rare,
                                   A.rare_cap, A.rare_count);
      if (A.want_grad && j >= 0 && (b2x != 0.0 || b2y != 0.0 || b2z != 0.0)) {
        atomicAdd(&A.gj[j], P.c_grad[0] * b2x);
        atomicAdd(&A.gj[(size_t)A.nj + j], P.c_grad[1] * b2y);
        atomicAdd(&A.gj[2 * (size_t)A.nj + j], P.c_grad[2] * b2z);
      }
      __syncwarp();
    }
    if (A.want_grad && ivalid) {
      atomicAdd(&A.gi[pi], -P.c_grad[0] * a1x);
      atomicAdd(&A.gi[(size_t)A.ni + pi], -P.c_grad[1] * a1y);
      atomicAdd(&A.gi[2 * (size_t)A.ni + pi], -P.c_grad[2] * a1z);
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) fsum += __shfl_xor_sync(0xffffffffu, fsum, d);
  if (lane == 0) red[warp] = fsum;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < WARPS; w++) s += red[w];
    A.partials[blockIdx.x] = s;
  }
}

void side_free(TileSide &S)
{
  cudaFree(S.d_key);
  cudaFree(S.d_key_sorted);
  cudaFree(S.d_val);
  cudaFree(S.d_sorted);
  cudaFree(S.d_cell_start);
  cudaFree(S.d_xs);
  cudaFree(S.d_ys);
  cudaFree(S.d_zs);
  cudaFree(S.d_gs);
  cudaFree(S.d_clusters);
  cudaFree(S.d_ncl);
  S = TileSide();
}

int side_setup(TileSide &S, int n)
{
  S.n = n;
  cudaError_t e;
  size_t const nb = sizeof(int) * (size_t)std::max(n, 1);
  if ((e = cudaMalloc(&S.d_key, nb)) != cudaSuccess) return (int)e;
  if ((e = cudaMalloc(&S.d_key_sorted, nb)) != cudaSuccess) return (int)e;
  if ((e = cudaMalloc(&S.d_val, nb)) != cudaSuccess) return (int)e;
  if ((e = cudaMalloc(&S.d_sorted, nb)) != cudaSuccess) return (int)e;
  size_t const nbd = sizeof(double) * (size_t)std::max(n, 1);
  if ((e = cudaMalloc(&S.d_xs, nbd)) != cudaSuccess) return (int)e;
  if ((e = cudaMalloc(&S.d_ys, nbd)) != cudaSuccess) return (int)e;
  if ((e = cudaMalloc(&S.d_zs, nbd)) != cudaSuccess) return (int)e;
  if ((e = cudaMalloc(&S.d_gs, 3 * nbd)) != cudaSuccess) return (int)e;
  if ((e = cudaMalloc(&S.d_cell_start, sizeof(int) * ((size_t)TILE_MAX_CELLS + 2))) != cudaSuccess) {
    return (int)e;
  }
  S.max_clusters = n / 32 + TILE_MAX_DIM * TILE_MAX_DIM + 1;
  if ((e = cudaMalloc(&S.d_clusters, sizeof(int4) * (size_t)S.max_clusters)) != cudaSuccess) {
    return (int)e;
  }
  if ((e = cudaMalloc(&S.d_ncl, sizeof(int))) != cudaSuccess) return (int)e;
  if ((e = cudaMemset(S.d_ncl, 0, sizeof(int))) != cudaSuccess) return (int)e;
  return 0;
}

int sort_side(TileBuffers &B, TileSide &S, const double *x, const double *y, const double *z,
              const int *gate, cudaStream_t stream)
{
  tile_key_kernel<<<(S.n + 255) / 256, 256, 0, stream>>>(x, y, z, S.n, B.d_grid, S.d_key, S.d_val,
                                                         gate);
  note_launch();
  size_t bytes = B.temp_bytes;
  cudaError_t e = cub::DeviceRadixSort::SortPairs(B.d_temp, bytes, S.d_key, S.d_key_sorted,
                                                  S.d_val, S.d_sorted, S.n, 0, TILE_KEY_BITS,
                                                  stream);
  if (e != cudaSuccess) return (int)e;
  note_launch();
  tile_cell_start_kernel<<<(TILE_MAX_CELLS + 1 + 255) / 256, 256, 0, stream>>>(
      S.d_key_sorted, S.n, B.d_grid, S.d_cell_start, gate);
  note_launch();
  return (int)cudaGetLastError();
}

}  // namespace

int tiles_setup(TileBuffers &B, int n1, int n2)
{
  bool const self = n1 <= 0;
  if (B.d_grid && B.side[1].n == n2 && B.side[0].n == (self ? 0 : n1)) return 0;
  tiles_free(B);
  B.self = self;
  cudaError_t e;
  if ((e = cudaMalloc(&B.d_grid, sizeof(TileGrid))) != cudaSuccess) return (int)e;
  if ((e = cudaMalloc(&B.d_counts, 8 * sizeof(unsigned))) != cudaSuccess) return (int)e;
  if ((e = cudaMemset(B.d_counts, 0, 8 * sizeof(unsigned))) != cudaSuccess) return (int)e;
  int rc = side_setup(B.side[1], n2);
  if (rc) return rc;
  if (!self) {
    rc = side_setup(B.side[0], n1);
    if (rc) return rc;
  }
  int const nmax = std::max(n1, n2);
  B.temp_bytes = 0;
  e = cub::DeviceRadixSort::SortPairs(nullptr, B.temp_bytes, B.side[1].d_key,
                                      B.side[1].d_key_sorted, B.side[1].d_val, B.side[1].d_sorted,
                                      nmax, 0, TILE_KEY_BITS);
  if (e != cudaSuccess) return (int)e;
  if ((e = cudaMalloc(&B.d_temp, std::max<size_t>(B.temp_bytes, 16))) != cudaSuccess) return (int)e;
  return 0;
}

void tiles_free(TileBuffers &B)
{
  cudaFree(B.d_grid);
  cudaFree(B.d_counts);
  cudaFree(B.d_temp);
  cudaFree(B.d_unit_hdr);
  cudaFree(B.d_tile_j);
  cudaFree(B.d_tile_m);
  side_free(B.side[0]);
  side_free(B.side[1]);
  B = TileBuffers();
}

int tiles_reserve(TileBuffers &B, unsigned units)
{
  if (units <= B.cap_units && B.d_unit_hdr) return 0;
  cudaFree(B.d_unit_hdr);
  cudaFree(B.d_tile_j);
  cudaFree(B.d_tile_m);
  B.d_unit_hdr = nullptr;
  B.d_tile_j = nullptr;
  B.d_tile_m = nullptr;
  B.cap_units = 0;
  cudaError_t e;
  size_t const slots = (size_t)units * TILE_UNIT * 32;
  if ((e = cudaMalloc(&B.d_unit_hdr, sizeof(int4) * (size_t)units)) != cudaSuccess) return (int)e;
  if ((e = cudaMalloc(&B.d_tile_j, sizeof(int) * slots)) != cudaSuccess) return (int)e;
  if ((e = cudaMalloc(&B.d_tile_m, sizeof(unsigned) * slots)) != cudaSuccess) return (int)e;
  B.cap_units = units;
  return 0;
}

int tiles_sort(TileBuffers &B, const double *x1, const double *y1, const double *z1,
               const double *x2, const double *y2, const double *z2, const int *gate,
               cudaStream_t stream)
{
  TileSide &J = B.side[1];
  tile_grid_kernel<<<1, GRID_THREADS, 0, stream>>>(x2, y2, z2, J.n, B.d_grid, gate);
  note_launch();
  int rc = sort_side(B, J, x2, y2, z2, gate, stream);
  if (rc) return rc;
  TileSide &I = B.self ? J : B.side[0];
  if (!B.self) {
    rc = sort_side(B, I, x1, y1, z1, gate, stream);
    if (rc) return rc;
  }
  tile_cluster_kernel<<<1, 1024, 0, stream>>>(I.d_cell_start, B.d_grid, I.d_clusters, I.d_ncl,
                                              I.max_clusters, gate);
  note_launch();
  return (int)cudaGetLastError();
}

int tiles_gather(TileBuffers &B, const double *x1, const double *y1, const double *z1,
                 const double *x2, const double *y2, const double *z2, const int *gate,
                 int gate_on, cudaStream_t stream)
{
  TileSide &I = B.side[0], &J = B.side[1];
  int const b1 = B.self ? 0 : (I.n + 255) / 256, b2 = (J.n + 255) / 256;
  tile_gather_kernel<<<b1 + b2, 256, 0, stream>>>(x1, y1, z1, I.d_sorted, B.self ? 0 : I.n, I.d_xs,
                                                  I.d_ys, I.d_zs, I.d_gs, x2, y2, z2, J.d_sorted,
                                                  J.n, J.d_xs, J.d_ys, J.d_zs, J.d_gs, b1, gate,
                                                  gate_on);
  note_launch();
  return (int)cudaGetLastError();
}

int tiles_scatter_add(TileBuffers &B, double *g1, size_t s1, double *g2, size_t s2,
                      const int *gate, int gate_on, cudaStream_t stream)
{
  TileSide &I = B.side[0], &J = B.side[1];
  int const b1 = B.self ? 0 : (I.n + 255) / 256, b2 = (J.n + 255) / 256;
  tile_scatter_add_kernel<<<b1 + b2, 256, 0, stream>>>(I.d_gs, I.d_sorted, B.self ? 0 : I.n, g1, s1,
                                                       J.d_gs, J.d_sorted, J.n, g2, s2, b1, gate,
                                                       gate_on);
  note_launch();
  return (int)cudaGetLastError();
}

int launch_tile_build(const TileBuildArgs &args, int exp_special, bool aniso, cudaStream_t stream)
{
  if (exp_special == 3) {
    if (aniso) tile_build_kernel<3, true><<<TILE_BLOCKS, TILE_THREADS, 0, stream>>>(args);
    else tile_build_kernel<3, false><<<TILE_BLOCKS, TILE_THREADS, 0, stream>>>(args);
  } else {
    if (aniso) tile_build_kernel<0, true><<<TILE_BLOCKS, TILE_THREADS, 0, stream>>>(args);
    else tile_build_kernel<0, false><<<TILE_BLOCKS, TILE_THREADS, 0, stream>>>(args);
  }
  note_launch();
  return (int)cudaGetLastError();
}

int launch_tile_walk(const TileWalkArgs &args, int exp_special, bool aniso, cudaStream_t stream)
{
  if (exp_special == 3) {
    if (aniso) tile_walk_kernel<3, true><<<TILE_BLOCKS, TILE_THREADS, 0, stream>>>(args);
    else tile_walk_kernel<3, false><<<TILE_BLOCKS, TILE_THREADS, 0, stream>>>(args);
  } else {
    if (aniso) tile_walk_kernel<0, true><<<TILE_BLOCKS, TILE_THREADS, 0, stream>>>(args);
    else tile_walk_kernel<0, false><<<TILE_BLOCKS, TILE_THREADS, 0, stream>>>(args);
  }
  note_launch();
  return (int)cudaGetLastError();
}

}  // namespace b200cv
