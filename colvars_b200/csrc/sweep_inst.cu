// sweep_inst.cu -- instantiations and dispatch of the tile sweep (sweep.cuh).
#include "kernels.cuh"

#include <atomic>

#ifndef B200CV_GEN_MINB
#define B200CV_GEN_MINB SWEEP_MINB
#endif

namespace b200cv {

// ============================ sweep dispatch =================================================

bool sweep_has_special(int en, int ed) { return en == 6 && ed == 12; }

template <int EXP, bool ANISO, int PBC, bool PL>
static int launch_sweep_t(const SweepArgs &args, int nitems, cudaStream_t stream)
{
  // general cells: the image search wants its registers less than the other sweeps want theirs
  constexpr int MINB = PBC == BC_GENERAL ? B200CV_GEN_MINB : SWEEP_MINB;
  auto kern = sweep_kernel<EXP, ANISO, PBC, PL, SWEEP_R, SWEEP_W, MINB>;
  // the opt-in shared-memory size is a per-device function attribute
  static std::atomic<unsigned long long> attr_done{0ull};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return (int)e;
  unsigned long long const bit = 1ull << (dev & 63);
  if (!(attr_done.load(std::memory_order_acquire) & bit)) {
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)sweep_smem_bytes());
    if (e != cudaSuccess) return (int)e;
    attr_done.fetch_or(bit, std::memory_order_release);
  }
  kern<<<nitems, SWEEP_W * 32, sweep_smem_bytes(), stream>>>(args);
  note_launch();
  return (int)cudaGetLastError();
}

template <int EXP, bool ANISO>
static int launch_sweep_ea(const SweepArgs &a, int n, int pbc, bool pl, cudaStream_t s)
{
  if (pbc == BC_NONE) {
    return pl ? launch_sweep_t<EXP, ANISO, BC_NONE, true>(a, n, s)
              : launch_sweep_t<EXP, ANISO, BC_NONE, false>(a, n, s);
  }
  if (pbc == BC_ORTHO) {
    return pl ? launch_sweep_t<EXP, ANISO, BC_ORTHO, true>(a, n, s)
              : launch_sweep_t<EXP, ANISO, BC_ORTHO, false>(a, n, s);
  }
  return pl ? launch_sweep_t<EXP, ANISO, BC_GENERAL, true>(a, n, s)
            : launch_sweep_t<EXP, ANISO, BC_GENERAL, false>(a, n, s);
}

template <int EXP>
static int launch_sweep_dinv(const SweepArgs &a, int n, int pbc, cudaStream_t s)
{
  if (pbc == BC_NONE) return launch_sweep_t<EXP, false, BC_NONE, false>(a, n, s);
  if (pbc == BC_ORTHO) return launch_sweep_t<EXP, false, BC_ORTHO, false>(a, n, s);
  return launch_sweep_t<EXP, false, BC_GENERAL, false>(a, n, s);
}

int launch_sweep(const SweepArgs &args, int nitems, int exp_special, bool aniso, int pbc,
                 bool pairlist, cudaStream_t stream)
{
  if (nitems <= 0) return 0;
  if (exp_special < 0) {
    // distanceInv: exponent 6 (the default) compiled in, any other even exponent at run time
    if (exp_special == -3) return launch_sweep_dinv<-3>(args, nitems, pbc, stream);
    return launch_sweep_dinv<-1>(args, nitems, pbc, stream);
  }
  if (exp_special == 3) {
    return aniso ? launch_sweep_ea<3, true>(args, nitems, pbc, pairlist, stream)
                 : launch_sweep_ea<3, false>(args, nitems, pbc, pairlist, stream);
  }
  return aniso ? launch_sweep_ea<0, true>(args, nitems, pbc, pairlist, stream)
               : launch_sweep_ea<0, false>(args, nitems, pbc, pairlist, stream);
}

#if defined(B200CV_TRACE)
int trace_dump(unsigned long long *out, int n)
{
  return (int)cudaMemcpyFromSymbol(out, g_trace, sizeof(unsigned long long) * 6 * (size_t)n);
}
#endif

}  // namespace b200cv
