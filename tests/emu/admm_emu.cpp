// tests/emu/admm_emu.cpp -- TEST INFRASTRUCTURE: host build of the warp-per-instance kernels of one generated family
// (matrix-parameter kernel, tail kernel, backward kernel) on top of the SIMT emulator tests/emu/simt/cuda_runtime.h.
// Compiled by tests/test_simt_emulation.py with g++ against a generated code directory:
//   -DCPG_SIMT_HOST_EMU -I tests/emu/simt -I <code_dir>/c/include -I <code_dir>/c/solver_code   + <code_dir>/c/src/cpg_blob.c
// The kernels are the product sources, unchanged; only the launch and the device memory are replaced by host calls.
#include "cuda_runtime.h"
// these kernels have no block-scope __shared__ state besides the (unused) mbarrier word: bind `extern __shared__ smem[]` to a
// plain host array (the interior-point build, ipm_simt.cpp, needs the thread_local form for its `__shared__ int next_inst`)
#undef __shared__
#define __shared__

#include "cpg_family.h"
#include "cpg_b200.h"
#include "cpg_blob_layout.h"
#include "admm_multi_kernel.cuh"
#include "grad_kernel.cuh"
#if CPG_FAM_MATPAR
#include "matpar_kernel.cuh"
#endif

extern "C" const unsigned long long CPG_B200_FN(cpg_blob_words)[];
extern "C" const unsigned long long CPG_B200_FN(cpg_tail_blob_words)[];
extern "C" const unsigned long long CPG_B200_FN(cpg_cblob_words)[];
extern "C" const unsigned long long CPG_B200_FN(cpg_gblob_words)[];
extern "C" const unsigned long long CPG_B200_FN(cpg_gS0_words)[];
extern "C" const unsigned long long CPG_B200_FN(cpg_mblob_words)[];
extern "C" const unsigned long long CPG_B200_FN(cpg_dblob_words)[];

namespace cpgb200 { alignas(128) uint8_t smem[256 * 1024]; }

namespace {
struct Fam {
  static constexpr int N = CPG_FAM_N, M = CPG_FAM_M;
  static constexpr int TRAIL = CPG_FAM_TRAIL_TILES, WARPS = CPG_FAM_WARPS, NI = CPG_FAM_NI, MULTI_STRIDE = CPG_FAM_MULTI_STRIDE;
  static constexpr int BLOB_BYTES_PAD = CPG_FAM_BLOB_BYTES_PAD;
  static constexpr int CBLOB_BYTES_PAD = CPG_FAM_CBLOB_BYTES_PAD;
  static constexpr int MAXREG = 255;
  static constexpr int W_STRIDE = CPG_FAM_W_STRIDE, S_STRIDE = CPG_FAM_S_STRIDE;
  static constexpr int TAIL_WARPS = CPG_FAM_TAIL_WARPS;
  static constexpr bool TAIL_STAGE = CPG_FAM_TAIL_STAGE != 0;
  static constexpr int GBLOB_BYTES_PAD = CPG_FAM_GBLOB_BYTES_PAD;
  static constexpr int GRAD_WARPS = CPG_FAM_GRAD_WARPS, GRAD_STRIDE = CPG_FAM_GRAD_STRIDE;
  static constexpr int DM_GROUPS = CPG_FAM_DM_GROUPS, DBLOB_BYTES_PAD = CPG_FAM_DBLOB_BYTES_PAD;
  static constexpr int DM_W8 = CPG_FAM_DM_W8, DM_STAGE = CPG_FAM_DM_STAGE, DM_BV = CPG_FAM_DM_BV;
#if CPG_FAM_MATPAR
  static constexpr int MAT_WARPS = CPG_FAM_MAT_WARPS;
  static constexpr int MAT_A_STRIDE = CPG_FAM_MAT_A_STRIDE, MAT_P_STRIDE = CPG_FAM_MAT_P_STRIDE;
  static constexpr int MAT_STRIDE = CPG_FAM_MAT_STRIDE, MAT_G_STRIDE = CPG_FAM_MAT_G_STRIDE;
#endif
};

cpgb200::Settings default_settings(int adaptive_rho_interval, double eps) {
  cpgb200::Settings st;
  st.max_iter = 4000; st.check_termination = 25; st.scaled_termination = 0; st.warm_start = 0;
  st.adaptive_rho = 1; st.adaptive_rho_interval = adaptive_rho_interval > 0 ? adaptive_rho_interval : 100;
  st.scaling = CPG_FAM_SCALING; st.pad = 0;
  st.eps_abs = eps; st.eps_rel = eps; st.eps_prim_inf = 1e-4; st.eps_dual_inf = 1e-4; st.alpha = 1.6; st.adaptive_rho_tolerance = 5.0;
  return st;
}
}  // namespace

extern "C" {

void emu_set_schedule(int mode) { simt::rt().schedule = mode; }
// synchronisation points executed since the last call, summed over all lanes: shuffles (incl. the ones inside __any_sync),
// __syncwarp, __syncthreads -- each is one step of a warp's dependent chain
void emu_counters(long long* out) {
  simt::Runtime& r = simt::rt();
  out[0] = r.n_exchange; out[1] = r.n_syncwarp; out[2] = r.n_syncthreads;
  r.n_exchange = r.n_syncwarp = r.n_syncthreads = 0;
}

int emu_dims(int* out) {   // n, m, npb, n_prim, n_dual, matpar
  const CpgBlobHeader* H = reinterpret_cast<const CpgBlobHeader*>(CPG_B200_FN(cpg_blob_words));
  out[0] = H->n; out[1] = H->m; out[2] = H->npb; out[3] = H->n_prim; out[4] = H->n_dual; out[5] = CPG_FAM_MATPAR;
  return 0;
}

// admm_matpar_kernel on `grid` blocks: the whole per-instance path (canonicalise P / A, equilibrate, assemble, factor, ADMM)
int emu_matpar_solve(int B, const double* params, double* prim, double* dual, double* sol_x, double* sol_y, double* obj,
                     int* iter, int* status, double* pri, double* dua, int grid, int adaptive_rho_interval, double eps) {
#if CPG_FAM_MATPAR
  cpgb200::BatchIO io;
  memset(&io, 0, sizeof(io));
  unsigned counter = 0;
  io.params = params; io.prim = prim; io.dual = dual; io.sol_x = sol_x; io.sol_y = sol_y; io.obj_val = obj; io.iter = iter;
  io.status = status; io.pri_res = pri; io.dua_res = dua; io.work_counter = &counter; io.B = B;
  const cpgb200::Settings st = default_settings(adaptive_rho_interval, eps);
  std::vector<double> scratch((size_t)grid * Fam::MAT_WARPS * Fam::MAT_G_STRIDE);
  simt::launch(grid, Fam::MAT_WARPS * 32, [&] {
    cpgb200::admm_matpar_kernel<Fam>(reinterpret_cast<const uint8_t*>(CPG_B200_FN(cpg_cblob_words)),
                                     reinterpret_cast<const uint8_t*>(CPG_B200_FN(cpg_tail_blob_words)),
                                     reinterpret_cast<const uint8_t*>(CPG_B200_FN(cpg_mblob_words)), scratch.data(), io, st);
  });
  return 0;
#else
  return 1;
#endif
}

// The two launches of cpg_solve_batch_device for shared-matrix families: admm_multi_kernel (NI instances per warp, generated
// straight-line KKT solve, lockstep CTA) and then admm_tail_kernel on whatever it handed off (rho updates, type changes)
static const double* g_warm_x0 = nullptr;     // set by emu_main_solve_warm for the duration of one solve
static const double* g_warm_y0 = nullptr;
int emu_main_solve(int B, const double* params, double* prim, double* dual, double* sol_x, double* sol_y, double* obj,
                   int* iter, int* status, double* pri, double* dua, int grid, int adaptive_rho_interval, double eps) {
  constexpr int WORDS = Fam::N + 2 * Fam::M + 2;
  std::vector<double> state((size_t)B * WORDS, 0.0);
  std::vector<int> ids(B);
  int count = 0;
  unsigned counter = 0;
  cpgb200::BatchIO io;
  memset(&io, 0, sizeof(io));
  io.params = params; io.prim = prim; io.dual = dual; io.sol_x = sol_x; io.sol_y = sol_y; io.obj_val = obj; io.iter = iter;
  io.status = status; io.pri_res = pri; io.dua_res = dua; io.B = B; io.work_counter = &counter;
  io.tail_count = &count; io.tail_ids = ids.data(); io.tail_state = state.data(); io.tail_capacity = B;
  cpgb200::Settings st = default_settings(adaptive_rho_interval, eps);
  std::vector<double> ws;
  if (g_warm_x0 && g_warm_y0) {          // warm start (osqp_warm_start): start points of the main kernel live in io.ws
    io.x0 = g_warm_x0; io.y0 = g_warm_y0; st.warm_start = 1;
    ws.assign((size_t)B * 2 * (Fam::M > 0 ? Fam::M : 1), 0.0);
    io.ws = ws.data();
  }
#if CPG_FAM_BIG        // schedule larger than shared memory: every instance is queued for the per-instance-factor kernel
  simt::launch((B + 255) / 256, 256, [&] { cpgb200::queue_all_kernel(io, WORDS, Fam::N + 2 * Fam::M,
                                                                        reinterpret_cast<const CpgBlobHeader*>(CPG_B200_FN(cpg_blob_words))->rho); });
#elif CPG_FAM_DMMA     // the family's main kernel is the tensor-core variant (groups of four warps, DMMA emulated lane by lane)
  simt::launch(grid, Fam::DM_GROUPS * 128, [&] {
    cpgb200::admm_dmma_kernel<Fam>(reinterpret_cast<const uint8_t*>(CPG_B200_FN(cpg_cblob_words)),
                                   reinterpret_cast<const uint8_t*>(CPG_B200_FN(cpg_dblob_words)), io, st);
  });
#else
  simt::launch(grid, Fam::WARPS * 32, [&] {
    cpgb200::admm_multi_kernel<Fam>(reinterpret_cast<const uint8_t*>(CPG_B200_FN(cpg_blob_words)), io, st);
  });
#endif
  const int handed_off = count;
  simt::launch(grid, Fam::TAIL_WARPS * 32, [&] {
    cpgb200::admm_tail_kernel<Fam>(reinterpret_cast<const uint8_t*>(CPG_B200_FN(cpg_cblob_words)),
                                   reinterpret_cast<const uint8_t*>(CPG_B200_FN(cpg_tail_blob_words)), io, st);
  });
  return handed_off;
}

int emu_main_solve_warm(int B, const double* params, const double* x0, const double* y0, double* prim, double* dual, double* sol_x,
                        double* sol_y, double* obj, int* iter, int* status, double* pri, double* dua, int grid,
                        int adaptive_rho_interval, double eps) {
  g_warm_x0 = x0; g_warm_y0 = y0;
  const int rc = emu_main_solve(B, params, prim, dual, sol_x, sol_y, obj, iter, status, pri, dua, grid, adaptive_rho_interval, eps);
  g_warm_x0 = g_warm_y0 = nullptr;
  return rc;
}

// admm_tail_kernel: every instance is queued at iteration 0 with the family's rho (the route an instance whose bounds
// change a constraint type takes), so the kernel factors K itself and runs the complete ADMM loop on its own factor
int emu_tail_solve(int B, const double* params, double* prim, double* dual, double* sol_x, double* sol_y, double* obj,
                   int* iter, int* status, double* pri, double* dua, int grid, int adaptive_rho_interval, double eps) {
  const CpgBlobHeader* H = reinterpret_cast<const CpgBlobHeader*>(CPG_B200_FN(cpg_blob_words));
  constexpr int WORDS = Fam::N + 2 * Fam::M + 2;
  std::vector<double> state((size_t)B * WORDS, 0.0);
  std::vector<int> ids(B);
  for (int b = 0; b < B; ++b) { ids[b] = b; state[(size_t)b * WORDS + Fam::N + 2 * Fam::M] = H->rho; }
  int count = B;
  cpgb200::BatchIO io;
  memset(&io, 0, sizeof(io));
  io.params = params; io.prim = prim; io.dual = dual; io.sol_x = sol_x; io.sol_y = sol_y; io.obj_val = obj; io.iter = iter;
  io.status = status; io.pri_res = pri; io.dua_res = dua; io.B = B;
  io.tail_count = &count; io.tail_ids = ids.data(); io.tail_state = state.data(); io.tail_capacity = B;
  const cpgb200::Settings st = default_settings(adaptive_rho_interval, eps);
  simt::launch(grid, Fam::TAIL_WARPS * 32, [&] {
    cpgb200::admm_tail_kernel<Fam>(reinterpret_cast<const uint8_t*>(CPG_B200_FN(cpg_cblob_words)),
                                   reinterpret_cast<const uint8_t*>(CPG_B200_FN(cpg_tail_blob_words)), io, st);
  });
  return 0;
}

// qp_grad_kernel (shared matrices) / qp_grad_kernel<Fam, true> (per-instance matrices)
int emu_gradient(int B, const double* params, const double* sol_x, const double* sol_y, const double* dprim, double* dparams,
                 double* dq, double* dl, double* du, double* dP, double* dA, int grid) {
  cpgb200::GradIO io;
  memset(&io, 0, sizeof(io));
  io.sol_y = sol_y; io.dprim = dprim; io.dparams = dparams; io.dq = dq; io.dl = dl; io.du = du;
  io.S0 = reinterpret_cast<const double*>(CPG_B200_FN(cpg_gS0_words)); io.B = B;
  const uint8_t* gblob = reinterpret_cast<const uint8_t*>(CPG_B200_FN(cpg_gblob_words));
  const uint8_t* tblob = reinterpret_cast<const uint8_t*>(CPG_B200_FN(cpg_tail_blob_words));
#if CPG_FAM_MATPAR
  std::vector<double> scratch((size_t)grid * Fam::GRAD_WARPS * Fam::MAT_G_STRIDE);
  io.params = params; io.sol_x = sol_x; io.dP = dP; io.dA = dA; io.a_scratch = scratch.data();
  io.mblob = reinterpret_cast<const uint8_t*>(CPG_B200_FN(cpg_mblob_words));
  simt::launch(grid, Fam::GRAD_WARPS * 32, [&] { cpgb200::qp_grad_kernel<Fam, true>(gblob, tblob, io); });
#else
  (void)params; (void)sol_x; (void)dP; (void)dA;
  simt::launch(grid, Fam::GRAD_WARPS * 32, [&] { cpgb200::qp_grad_kernel<Fam>(gblob, tblob, io); });
#endif
  return 0;
}

}  // extern "C"
