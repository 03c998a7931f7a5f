// keysort.cu -- see keysort.cuh.
#include "keysort.cuh"

#include <algorithm>

#include "kernels.cuh"

namespace b200cv {

namespace {

constexpr int SCAN_THREADS = 1024;
constexpr int SCAN_PER_THREAD = 4;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_PER_THREAD;

// exclusive scan inside a tile of SCAN_TILE cells; the tile's total goes to tile[blockIdx.x]
__global__ void __launch_bounds__(SCAN_THREADS)
keysort_scan_kernel(const int *__restrict__ cnt, int ncells, int *cell_start, int *tile,
                    const int *gate)
{
  __shared__ int wsum[SCAN_THREADS / 32];
  if (gate != nullptr && *gate != 1) return;
  int const base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_PER_THREAD;
  int v[SCAN_PER_THREAD], s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_PER_THREAD; k++) {
    v[k] = base + k < ncells ? cnt[base + k] : 0;
    s += v[k];
  }
  int const lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = s;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int const t = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += t;
  }
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = wsum[lane];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int const t = __shfl_up_sync(0xffffffffu, w, d);
      if (lane >= d) w += t;
    }
    wsum[lane] = w;  // inclusive over the warps
  }
  __syncthreads();
  int run = incl - s + (warp > 0 ? wsum[warp - 1] : 0);
#pragma unroll
  for (int k = 0; k < SCAN_PER_THREAD; k++) {
    if (base + k < ncells) cell_start[base + k] = run;
    run += v[k];
  }
  if (threadIdx.x == SCAN_THREADS - 1) tile[blockIdx.x] = run;
}

// + the totals of the tiles in front
__global__ void __launch_bounds__(SCAN_THREADS)
keysort_offset_kernel(int ncells, int *cell_start, const int *__restrict__ tile, const int *gate)
{
  __shared__ int offset;
  if (gate != nullptr && *gate != 1) return;
  if (blockIdx.x == 0) return;  // nothing in front of the first tile
  if (threadIdx.x < 32) {
    int s = 0;
    for (int t = threadIdx.x; t < (int)blockIdx.x; t += 32) s += tile[t];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    if (threadIdx.x == 0) offset = s;
  }
  __syncthreads();
  int const base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_PER_THREAD;
#pragma unroll
  for (int k = 0; k < SCAN_PER_THREAD; k++) {
    if (base + k < ncells) cell_start[base + k] += offset;
  }
}

__global__ void keysort_scatter_kernel(const int *__restrict__ key, int n, int fine_bits, int *cnt,
                                       const int *__restrict__ cell_start, int *tmp_key,
                                       int *tmp_val, const int *gate)
{
  int const i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (gate != nullptr && *gate != 1) return;
  int const k = key[i], c = k >> fine_bits;
  int const slot = cell_start[c] + atomicSub(&cnt[c], 1) - 1;
  tmp_key[slot] = k;
  tmp_val[slot] = i;
}

__global__ void keysort_rank_kernel(const int *__restrict__ key, int n, int fine_bits,
                                    const int *__restrict__ cell_start,
                                    const int *__restrict__ tmp_key,
                                    const int *__restrict__ tmp_val, int *sorted, const int *gate)
{
  int const i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (gate != nullptr && *gate != 1) return;
  int const k = key[i], c = k >> fine_bits;
  int const lo = cell_start[c], hi = cell_start[c + 1];
  int rank = 0;
  for (int e = lo; e < hi; e++) {
    int const ke = tmp_key[e], ve = tmp_val[e];
    rank += (ke < k || (ke == k && ve < i)) ? 1 : 0;
  }
  sorted[lo + rank] = i;
}

}  // namespace

int keysort_setup(KeySort &S, int n, int ncells_max)
{
  if (S.d_cnt && S.n == n && S.ncells_max == ncells_max) return 0;
  keysort_free(S);
  S.n = n;
  S.ncells_max = ncells_max;
  cudaError_t e;
  size_t const nc = (size_t)ncells_max + 2;
  if ((e = cudaMalloc(&S.d_cnt, sizeof(int) * nc)) != cudaSuccess) return (int)e;
  if ((e = cudaMemset(S.d_cnt, 0, sizeof(int) * nc)) != cudaSuccess) return (int)e;
  size_t const tiles = (nc + SCAN_TILE - 1) / SCAN_TILE;
  if ((e = cudaMalloc(&S.d_tile, sizeof(int) * tiles)) != cudaSuccess) return (int)e;
  size_t const nb = sizeof(int) * (size_t)std::max(n, 1);
  if ((e = cudaMalloc(&S.d_tmp_key, nb)) != cudaSuccess) return (int)e;
  if ((e = cudaMalloc(&S.d_tmp_val, nb)) != cudaSuccess) return (int)e;
  return 0;
}

void keysort_free(KeySort &S)
{
  cudaFree(S.d_cnt);
  cudaFree(S.d_tile);
  cudaFree(S.d_tmp_key);
  cudaFree(S.d_tmp_val);
  S = KeySort();
}

int keysort_launches(const KeySort &S)
{
  return (S.ncells_max + 2 + SCAN_TILE - 1) / SCAN_TILE > 1 ? 4 : 3;
}

int keysort_run(KeySort &S, const int *key, int fine_bits, int *sorted, int *cell_start,
                const int *gate, cudaStream_t stream)
{
  int const n = S.n, nc = S.ncells_max + 2;
  int const blocks = std::max(1, (n + 255) / 256), tiles = (nc + SCAN_TILE - 1) / SCAN_TILE;
  keysort_scan_kernel<<<tiles, SCAN_THREADS, 0, stream>>>(S.d_cnt, nc, cell_start, S.d_tile, gate);
  note_launch();
  if (tiles > 1) {
    keysort_offset_kernel<<<tiles, SCAN_THREADS, 0, stream>>>(nc, cell_start, S.d_tile, gate);
    note_launch();
  }
  keysort_scatter_kernel<<<blocks, 256, 0, stream>>>(key, n, fine_bits, S.d_cnt, cell_start,
                                                     S.d_tmp_key, S.d_tmp_val, gate);
  note_launch();
  keysort_rank_kernel<<<blocks, 256, 0, stream>>>(key, n, fine_bits, cell_start, S.d_tmp_key,
                                                  S.d_tmp_val, sorted, gate);
  note_launch();
  return (int)cudaGetLastError();
}

}  // namespace b200cv
