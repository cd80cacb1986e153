// tests/emu/ipm_simt.cpp -- TEST INFRASTRUCTURE: the interior-point kernel's DEVICE code path (ipm_kernel itself: 256-thread
// CTA, barrier-separated phases, gather plans finished with shuffles, REDUX max reductions, persistent CTA pulling instances
// from the work counter) compiled unchanged against the SIMT emulator tests/emu/simt/cuda_runtime.h.  tests/emu/ipm_emu.cpp
// is the older host build of the same source (-DCPG_IPM_HOST_EMU: phases serialised through host-only branches); this one
// executes what the GPU executes, thread by thread, and adds the thread-schedule race check.
#include "cuda_runtime.h"
#include <vector>
#include "cpg_ipm_family.h"
#include "ipm_kernel.cuh"

namespace cpgipm { alignas(128) thread_local unsigned char smem_raw[256 * 1024]; }

extern "C" void ipm_simt_set_schedule(int mode) { simt::rt().schedule = mode; }

extern "C" int ipm_simt_solve(const unsigned char* sblob, const unsigned char* gblob, int B, const double* params,
                              double* prim, double* dual, double* x, double* y, double* z, double* s,
                              double* obj, int* iter, int* status, double* pres, double* dres, int maxit, int grid) {
  using namespace cpgipm;
  static_assert(SMEM_BYTES <= 256 * 1024, "emulated shared memory too small");
  std::vector<double> best(size_t(grid) * BEST_STRIDE);
  int counter = 0;
  IpmSettings stg{maxit, 0, 1e-8, 1e-8, 1e-8, 1e-4, 5e-5, 5e-5};
  IpmIO io{B, params, prim, dual, x, y, z, s, obj, iter, status, pres, dres, best.data(), &counter};
  simt::launch(grid, CPG_IPM_THREADS, [&] { ipm_kernel(sblob, gblob, stg, io); });
  return 0;
}
