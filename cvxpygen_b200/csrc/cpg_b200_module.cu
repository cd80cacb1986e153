// cpg_b200_module.cu -- C-ABI runtime of one generated ADMM-CUDA solver library.
// Compiled once per problem family together with the generated cpg_family.h (compile-time sizes),
// cpg_blob_layout.h (blob header struct) and cpg_blob.c (the constants blob).
// Implements include/cpg_b200.h; see that header for the reference interfaces it stands beside.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <thread>
#include <vector>

#include "cpg_family.h"
#include "cpg_b200.h"
#include "cpg_blob_layout.h"
#include "admm_multi_kernel.cuh"
#include "grad_kernel.cuh"
#if CPG_FAM_MATPAR
#include "matpar_kernel.cuh"
#endif

extern "C" const unsigned long long CPG_B200_FN(cpg_blob_words)[];
extern "C" const unsigned int CPG_B200_FN(cpg_blob_nbytes);
extern "C" const unsigned long long CPG_B200_FN(cpg_tail_blob_words)[];
extern "C" const unsigned int CPG_B200_FN(cpg_tail_blob_nbytes);
extern "C" const unsigned long long CPG_B200_FN(cpg_cblob_words)[];     // main blob without the tile schedule (tail kernel)
extern "C" const unsigned int CPG_B200_FN(cpg_cblob_nbytes);
extern "C" const unsigned long long CPG_B200_FN(cpg_gblob_words)[];     // backward-pass constants (shared memory)
extern "C" const unsigned int CPG_B200_FN(cpg_gblob_nbytes);
extern "C" const unsigned long long CPG_B200_FN(cpg_gS0_words)[];       // regularised KKT values in slot order (global)
extern "C" const unsigned int CPG_B200_FN(cpg_gS0_nbytes);
extern "C" const unsigned long long CPG_B200_FN(cpg_mblob_words)[];     // matrix-parameter tables (global; 16 bytes of zeros when unused)
extern "C" const unsigned int CPG_B200_FN(cpg_mblob_nbytes);
extern "C" const unsigned long long CPG_B200_FN(cpg_dblob_words)[];     // tables of the FP64 tensor-core solve (32 bytes of zeros when unused)
extern "C" const unsigned int CPG_B200_FN(cpg_dblob_nbytes);

namespace {

struct Fam {
  static constexpr int N = CPG_FAM_N, M = CPG_FAM_M;
  static constexpr int TRAIL = CPG_FAM_TRAIL_TILES;
  static constexpr int WARPS = CPG_FAM_WARPS;
  static constexpr int BLOB_BYTES_PAD = CPG_FAM_BLOB_BYTES_PAD;
  static constexpr int CBLOB_BYTES_PAD = CPG_FAM_CBLOB_BYTES_PAD;
  // registers are allocated to warps four at a time: the file (65 536) is divided among ceil(WARPS / 4) * 4 warps
  static constexpr int WARPS_ALLOC = (CPG_FAM_WARPS + 3) / 4 * 4;
  static constexpr int MAXREG = (65536 / (WARPS_ALLOC * 32)) / 8 * 8 > 255 ? 255 : (65536 / (WARPS_ALLOC * 32)) / 8 * 8;
  static constexpr int W_STRIDE = CPG_FAM_W_STRIDE;       // doubles per warp work vector
  static constexpr int S_STRIDE = CPG_FAM_S_STRIDE;       // doubles per warp factor storage (tail kernel)
  static constexpr int TAIL_WARPS = CPG_FAM_TAIL_WARPS;
  static constexpr bool TAIL_STAGE = CPG_FAM_TAIL_STAGE != 0;   // the tail kernel stages the compact blob (else: reads it through L2)
  static constexpr int GBLOB_BYTES_PAD = CPG_FAM_GBLOB_BYTES_PAD;
  static constexpr int GRAD_WARPS = CPG_FAM_GRAD_WARPS;
  static constexpr int GRAD_STRIDE = CPG_FAM_GRAD_STRIDE;
  static constexpr int NI = CPG_FAM_NI;                     // instances per warp in the main kernel (2 or 4)
  static constexpr int MULTI_STRIDE = CPG_FAM_MULTI_STRIDE; // doubles per warp: interleaved work vectors + batched-row slots
  static constexpr int DM_GROUPS = CPG_FAM_DM_GROUPS;       // tensor-core main kernel: groups of four warps (eight instances) per CTA
  static constexpr int DBLOB_BYTES_PAD = CPG_FAM_DBLOB_BYTES_PAD;
  static constexpr int DM_W8 = CPG_FAM_DM_W8, DM_STAGE = CPG_FAM_DM_STAGE, DM_BV = CPG_FAM_DM_BV;   // doubles
#if CPG_FAM_MATPAR
  static constexpr int MAT_WARPS = CPG_FAM_MAT_WARPS;       // matrix-parameter kernel: warps per CTA
  static constexpr int MAT_A_STRIDE = CPG_FAM_MAT_A_STRIDE, MAT_P_STRIDE = CPG_FAM_MAT_P_STRIDE;
  static constexpr int MAT_STRIDE = CPG_FAM_MAT_STRIDE;     // doubles of shared memory per warp: w | S | Pv
  static constexpr int MAT_G_STRIDE = CPG_FAM_MAT_G_STRIDE; // doubles of global scratch per warp: Av | D Dinv E Einv
#endif
};
#if CPG_FAM_MATPAR
constexpr int MAT_SMEM_BYTES = Fam::MAT_WARPS * Fam::MAT_STRIDE * 8;
#endif
constexpr int SMEM_BYTES = Fam::BLOB_BYTES_PAD + Fam::WARPS * Fam::MULTI_STRIDE * 8;
constexpr int DMMA_SMEM_BYTES = Fam::CBLOB_BYTES_PAD + Fam::DBLOB_BYTES_PAD +
                                Fam::DM_GROUPS * ((Fam::DM_W8 + Fam::DM_STAGE) * 8 + 4 * Fam::DM_BV * 8 + 8) + 16;
constexpr int TAIL_SMEM_BYTES = (Fam::TAIL_STAGE ? Fam::CBLOB_BYTES_PAD : 0) + Fam::TAIL_WARPS * (Fam::W_STRIDE + Fam::S_STRIDE) * 8;
constexpr int TAIL_WORDS = Fam::N + 2 * Fam::M + 2;
constexpr int GRAD_SMEM_BYTES = Fam::GBLOB_BYTES_PAD + Fam::GRAD_WARPS * Fam::GRAD_STRIDE * 8;

struct Ctx {
  bool ready = false;
  int device = -1, n_sm = 0;
  uint8_t* d_blob = nullptr;
  uint8_t* d_tail_blob = nullptr;
  uint8_t* d_cblob = nullptr;
  uint8_t* d_gblob = nullptr;
  double* d_gS0 = nullptr;
  uint8_t* d_mblob = nullptr;
  uint8_t* d_dblob = nullptr;
  double* d_mat_scratch = nullptr;       // per-warp slices for the scaled entries of A (matrix-parameter kernel)
  int cap_G = 0;
  double *g_soly = nullptr, *g_dprim = nullptr, *g_dparams = nullptr, *g_dq = nullptr, *g_dl = nullptr, *g_du = nullptr;
  unsigned int* d_counter = nullptr;
  int* d_tail_count = nullptr;
  int* d_tail_ids = nullptr;
  double* d_tail_state = nullptr;
  int tail_cap = 0;
  // staging for the host-buffer entry point
  int cap_B = 0;
  double* d_ws = nullptr;      // warm-start scratch of the main kernel (BatchIO::ws)
  int cap_ws = 0;
  double *d_params = nullptr, *d_x0 = nullptr, *d_y0 = nullptr, *d_prim = nullptr, *d_dual = nullptr;
  double *d_solx = nullptr, *d_soly = nullptr, *d_obj = nullptr, *d_pri = nullptr, *d_dua = nullptr;
  int *d_iter = nullptr, *d_status = nullptr;
  int launches = 0;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};   // solve: before / after main / after tail; [3] unused
  cudaEvent_t gev[2] = {nullptr, nullptr};                   // backward pass: before / after
  bool ev_solve = false, ev_grad = false;
  char err[256] = {0};
};
// One context per DEVICE; every host thread works on the context it selected last (cpg_b200_init / cpg_b200_use_device), so
// a multi-GPU caller runs one host thread per device on one loaded library (cpg_solve_batch_host_multi does exactly that).
constexpr int MAX_DEVICES = 16;
Ctx ctxs[MAX_DEVICES];
thread_local Ctx* cur_ctx = &ctxs[0];
#define g (*cur_ctx)

#define CK(call)                                                                                  \
  do {                                                                                            \
    cudaError_t e_ = (call);                                                                      \
    if (e_ != cudaSuccess) {                                                                      \
      snprintf(g.err, sizeof(g.err), "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
      return CPG_B200_ERR_CUDA;                                                                   \
    }                                                                                             \
  } while (0)

// Every entry point runs on the device of the calling thread's context.
#define USE_DEVICE()                                                                              \
  do {                                                                                            \
    if (!g.ready) { snprintf(g.err, sizeof(g.err), "cpg_b200_init has not been called"); return CPG_B200_ERR_NOT_INIT; } \
    CK(cudaSetDevice(g.device));                                                                  \
  } while (0)

int ensure_tail(int B) {
  if (B <= g.tail_cap) return CPG_B200_OK;
  if (g.d_tail_ids) cudaFree(g.d_tail_ids);
  if (g.d_tail_state) cudaFree(g.d_tail_state);
  g.d_tail_ids = nullptr; g.d_tail_state = nullptr; g.tail_cap = 0;
  CK(cudaMalloc(&g.d_tail_ids, sizeof(int) * (size_t)B));
  CK(cudaMalloc(&g.d_tail_state, sizeof(double) * (size_t)B * TAIL_WORDS));
  g.tail_cap = B;
  return CPG_B200_OK;
}

template <class T>
int grow(T** p, size_t count) {
  if (*p) cudaFree(*p);
  *p = nullptr;
  CK(cudaMalloc(p, sizeof(T) * count));
  return CPG_B200_OK;
}

// device-side alias of a pinned (cudaHostAlloc / cudaHostRegister) host buffer under unified addressing, else null
template <class Tp> Tp* mapped_view(Tp* host) {
  if (!host) return nullptr;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, host) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return a.type == cudaMemoryTypeHost ? static_cast<Tp*>(a.devicePointer) : nullptr;
}

int ensure_staging(int B) {
  if (B <= g.cap_B) return CPG_B200_OK;
  int rc;
  g.cap_B = 0;                 // a failed allocation below must not leave a stale capacity behind
  const int np = reinterpret_cast<const CpgBlobHeader*>(CPG_B200_FN(cpg_blob_words))->npb;
  const int npr = reinterpret_cast<const CpgBlobHeader*>(CPG_B200_FN(cpg_blob_words))->n_prim;
  const int ndu = reinterpret_cast<const CpgBlobHeader*>(CPG_B200_FN(cpg_blob_words))->n_dual;
  if ((rc = grow(&g.d_params, (size_t)B * (np > 0 ? np : 1)))) return rc;
  if ((rc = grow(&g.d_x0, (size_t)B * Fam::N))) return rc;
  if ((rc = grow(&g.d_y0, (size_t)B * (Fam::M > 0 ? Fam::M : 1)))) return rc;
  if ((rc = grow(&g.d_prim, (size_t)B * (npr > 0 ? npr : 1)))) return rc;
  if ((rc = grow(&g.d_dual, (size_t)B * (ndu > 0 ? ndu : 1)))) return rc;
  if ((rc = grow(&g.d_solx, (size_t)B * Fam::N))) return rc;
  if ((rc = grow(&g.d_soly, (size_t)B * (Fam::M > 0 ? Fam::M : 1)))) return rc;
  if ((rc = grow(&g.d_obj, (size_t)B))) return rc;
  if ((rc = grow(&g.d_pri, (size_t)B))) return rc;
  if ((rc = grow(&g.d_dua, (size_t)B))) return rc;
  if ((rc = grow(&g.d_iter, (size_t)B))) return rc;
  if ((rc = grow(&g.d_status, (size_t)B))) return rc;
  g.cap_B = B;
  return CPG_B200_OK;
}

}  // namespace

extern "C" {

const char* CPG_B200_FN(cpg_b200_last_error)(void) { return g.err; }
int CPG_B200_FN(cpg_b200_launch_count)(void) { return g.launches; }

int CPG_B200_FN(cpg_b200_kernel_times)(float* main_ms, float* tail_ms, float* grad_ms) {
  USE_DEVICE();
  if (main_ms) *main_ms = -1.f;
  if (tail_ms) *tail_ms = -1.f;
  if (grad_ms) *grad_ms = -1.f;
  if (g.ev_solve) {
    CK(cudaEventSynchronize(g.ev[2]));
    if (main_ms) CK(cudaEventElapsedTime(main_ms, g.ev[0], g.ev[1]));
    if (tail_ms) CK(cudaEventElapsedTime(tail_ms, g.ev[1], g.ev[2]));
  }
  if (g.ev_grad) {
    CK(cudaEventSynchronize(g.gev[1]));
    if (grad_ms) CK(cudaEventElapsedTime(grad_ms, g.gev[0], g.gev[1]));
  }
  return CPG_B200_OK;
}

void CPG_B200_FN(cpg_b200_default_settings)(CpgB200Settings* s) {
  if (!s) return;
  s->max_iter = 4000; s->check_termination = 25; s->scaled_termination = 0; s->warm_start = 0;
  s->adaptive_rho = 1; s->adaptive_rho_interval = 0; s->scaling = CPG_FAM_SCALING; s->host_zero_copy = 1;
  s->eps_abs = 1e-3; s->eps_rel = 1e-3; s->eps_prim_inf = 1e-4; s->eps_dual_inf = 1e-4;
  s->alpha = 1.6; s->adaptive_rho_tolerance = 5.0;
}

int CPG_B200_FN(cpg_b200_dims)(CpgB200Dims* out) {
  if (!out) return CPG_B200_ERR_BAD_ARG;
  const CpgBlobHeader* H = reinterpret_cast<const CpgBlobHeader*>(CPG_B200_FN(cpg_blob_words));
  out->n_var = H->n; out->n_con = H->m; out->n_param = H->npb; out->n_prim = H->n_prim; out->n_dual = H->n_dual;
  out->blob_bytes = H->total_bytes;
#if CPG_FAM_DMMA
  out->warps_per_cta = Fam::DM_GROUPS * 4; out->smem_bytes = DMMA_SMEM_BYTES;
#else
  out->warps_per_cta = Fam::WARPS; out->smem_bytes = SMEM_BYTES;
#endif
  return CPG_B200_OK;
}

int CPG_B200_FN(cpg_b200_load_constants)(const void* blob, int nbytes) {
  USE_DEVICE();
  if (!blob || nbytes <= 0 || nbytes > Fam::BLOB_BYTES_PAD) return CPG_B200_ERR_BAD_ARG;
  const CpgBlobHeader* H = reinterpret_cast<const CpgBlobHeader*>(blob);
  if (H->n != Fam::N || H->m != Fam::M || (int)H->total_bytes != nbytes || H->n_trail_tiles > Fam::TRAIL)
    return CPG_B200_ERR_BAD_ARG;
  CK(cudaMemcpy(g.d_blob, blob, nbytes, cudaMemcpyHostToDevice));
  return CPG_B200_OK;
}

int CPG_B200_FN(cpg_b200_load_constants_all)(const void* blob, int nbytes, const void* cblob, int cnbytes,
                                             const void* tail_blob, int tnbytes, const void* gblob, int gnbytes,
                                             const void* gS0, int snbytes) {
  USE_DEVICE();
  if (!blob || !cblob || !tail_blob || !gblob || !gS0) return CPG_B200_ERR_BAD_ARG;
  if (nbytes != (int)CPG_B200_FN(cpg_blob_nbytes) || cnbytes != (int)CPG_B200_FN(cpg_cblob_nbytes) ||
      tnbytes != (int)CPG_B200_FN(cpg_tail_blob_nbytes) || gnbytes != (int)CPG_B200_FN(cpg_gblob_nbytes) ||
      snbytes != (int)CPG_B200_FN(cpg_gS0_nbytes))
    return CPG_B200_ERR_BAD_ARG;          // a different layout needs a regenerated library
  {
    // the tile headers of the per-instance triangular solves are compiled into this library (constant memory) and the entry words'
    // format is a compile-time choice: the new tables must carry the same ones
    const CpgTailHeader* th = reinterpret_cast<const CpgTailHeader*>(tail_blob);
    const CpgTailHeader* th0 = reinterpret_cast<const CpgTailHeader*>(CPG_B200_FN(cpg_tail_blob_words));
    const size_t nt = (size_t)th0->n_fwd_tiles + th0->n_bwd_tiles;
    if (th->pad0 != th0->pad0 || th->n_fwd_tiles != th0->n_fwd_tiles || th->n_bwd_tiles != th0->n_bwd_tiles || th->off_i32 != th0->off_i32 ||
        th->i_tiles != th0->i_tiles ||
        memcmp(reinterpret_cast<const char*>(tail_blob) + th->off_i32 + 4 * (size_t)th->i_tiles,
               reinterpret_cast<const char*>(CPG_B200_FN(cpg_tail_blob_words)) + th0->off_i32 + 4 * (size_t)th0->i_tiles, nt * 32) != 0) {
      snprintf(g.err, sizeof(g.err), "the new tables have another solve structure than the one compiled into this library: regenerate it");
      return CPG_B200_ERR_BAD_ARG;
    }
  }
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(g.d_blob, blob, nbytes, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(g.d_cblob, cblob, cnbytes, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(g.d_tail_blob, tail_blob, tnbytes, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(g.d_gblob, gblob, gnbytes, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(g.d_gS0, gS0, snbytes, cudaMemcpyHostToDevice));
  return CPG_B200_OK;
}

/* tables of the tensor-core solve after a shared-parameter update (same sizes: the sparsity structure is unchanged) */
int CPG_B200_FN(cpg_b200_load_dmma_constants)(const void* dblob, int nbytes) {
  USE_DEVICE();
#if CPG_FAM_DMMA
  if (!dblob || nbytes != (int)CPG_B200_FN(cpg_dblob_nbytes)) return CPG_B200_ERR_BAD_ARG;
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(g.d_dblob, dblob, nbytes, cudaMemcpyHostToDevice));
  return CPG_B200_OK;
#else
  (void)dblob; (void)nbytes;
  return CPG_B200_OK;           // this library runs the straight-line schedule of the main blob
#endif
}

int CPG_B200_FN(cpg_b200_load_mat_constants)(const void* mblob, int nbytes) {
  USE_DEVICE();
#if CPG_FAM_MATPAR
  if (!mblob || nbytes != (int)CPG_B200_FN(cpg_mblob_nbytes)) return CPG_B200_ERR_BAD_ARG;
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(g.d_mblob, mblob, nbytes, cudaMemcpyHostToDevice));
  return CPG_B200_OK;
#else
  (void)mblob; (void)nbytes;
  snprintf(g.err, sizeof(g.err), "this library was generated without per-instance matrix parameters");
  return CPG_B200_ERR_BAD_ARG;
#endif
}

int CPG_B200_FN(cpg_b200_use_device)(int device) {
  if (device < 0 || device >= MAX_DEVICES) return CPG_B200_ERR_BAD_ARG;
  if (!ctxs[device].ready) {
    snprintf(g.err, sizeof(g.err), "cpg_b200_init(%d) has not been called", device);
    return CPG_B200_ERR_NOT_INIT;
  }
  cur_ctx = &ctxs[device];
  return CPG_B200_OK;
}

int CPG_B200_FN(cpg_b200_init)(int device) {
  if (device < 0 || device >= MAX_DEVICES) {
    snprintf(g.err, sizeof(g.err), "device index %d outside [0, %d)", device, MAX_DEVICES);
    return CPG_B200_ERR_BAD_ARG;
  }
  cur_ctx = &ctxs[device];              // this thread now works on the context of `device`
  g.err[0] = 0;
  CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) {
    snprintf(g.err, sizeof(g.err), "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    return CPG_B200_ERR_CUDA;
  }
  g.device = device; g.n_sm = prop.multiProcessorCount;
  if (!g.d_blob) CK(cudaMalloc(&g.d_blob, Fam::BLOB_BYTES_PAD));
  CK(cudaMemcpy(g.d_blob, CPG_B200_FN(cpg_blob_words), CPG_B200_FN(cpg_blob_nbytes), cudaMemcpyHostToDevice));
  if (!g.d_cblob) CK(cudaMalloc(&g.d_cblob, Fam::CBLOB_BYTES_PAD));
  CK(cudaMemcpy(g.d_cblob, CPG_B200_FN(cpg_cblob_words), CPG_B200_FN(cpg_cblob_nbytes), cudaMemcpyHostToDevice));
  if (!g.d_gblob) CK(cudaMalloc(&g.d_gblob, Fam::GBLOB_BYTES_PAD));
  CK(cudaMemcpy(g.d_gblob, CPG_B200_FN(cpg_gblob_words), CPG_B200_FN(cpg_gblob_nbytes), cudaMemcpyHostToDevice));
  if (!g.d_gS0) CK(cudaMalloc(&g.d_gS0, CPG_B200_FN(cpg_gS0_nbytes)));
  CK(cudaMemcpy(g.d_gS0, CPG_B200_FN(cpg_gS0_words), CPG_B200_FN(cpg_gS0_nbytes), cudaMemcpyHostToDevice));
#if CPG_FAM_GRAD
#if CPG_FAM_MATPAR
  CK(cudaFuncSetAttribute(cpgb200::qp_grad_kernel<Fam, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, GRAD_SMEM_BYTES));
#else
  CK(cudaFuncSetAttribute(cpgb200::qp_grad_kernel<Fam>, cudaFuncAttributeMaxDynamicSharedMemorySize, GRAD_SMEM_BYTES));
#endif
#endif
  if (!g.d_tail_blob) CK(cudaMalloc(&g.d_tail_blob, CPG_B200_FN(cpg_tail_blob_nbytes)));
  CK(cudaMemcpy(g.d_tail_blob, CPG_B200_FN(cpg_tail_blob_words), CPG_B200_FN(cpg_tail_blob_nbytes), cudaMemcpyHostToDevice));
#if !CPG_FAM_MATPAR
  CK(cudaFuncSetAttribute(cpgb200::admm_tail_kernel<Fam>, cudaFuncAttributeMaxDynamicSharedMemorySize, TAIL_SMEM_BYTES));
#endif
#if CPG_FAM_MATPAR
  if (!g.d_mblob) CK(cudaMalloc(&g.d_mblob, CPG_B200_FN(cpg_mblob_nbytes)));
  CK(cudaMemcpy(g.d_mblob, CPG_B200_FN(cpg_mblob_words), CPG_B200_FN(cpg_mblob_nbytes), cudaMemcpyHostToDevice));
  CK(cudaFuncSetAttribute(cpgb200::admm_matpar_kernel<Fam>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAT_SMEM_BYTES));
  if (!g.d_mat_scratch) CK(cudaMalloc(&g.d_mat_scratch, sizeof(double) * (size_t)g.n_sm * (Fam::MAT_WARPS > Fam::GRAD_WARPS ? Fam::MAT_WARPS : Fam::GRAD_WARPS) * Fam::MAT_G_STRIDE));
#endif
  if (!g.d_counter) CK(cudaMalloc(&g.d_counter, sizeof(unsigned int)));
  if (!g.d_tail_count) CK(cudaMalloc(&g.d_tail_count, sizeof(int)));
#if CPG_FAM_BIG || CPG_FAM_MATPAR
  // nothing to prepare: the main kernel is not part of this library
#elif CPG_FAM_DMMA
  if (!g.d_dblob) CK(cudaMalloc(&g.d_dblob, Fam::DBLOB_BYTES_PAD));
  CK(cudaMemcpy(g.d_dblob, CPG_B200_FN(cpg_dblob_words), CPG_B200_FN(cpg_dblob_nbytes), cudaMemcpyHostToDevice));
  CK(cudaFuncSetAttribute(cpgb200::admm_dmma_kernel<Fam>, cudaFuncAttributeMaxDynamicSharedMemorySize, DMMA_SMEM_BYTES));
#else
  CK(cudaFuncSetAttribute(cpgb200::admm_multi_kernel<Fam>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
#endif
  for (int k = 0; k < 4; ++k) if (!g.ev[k]) CK(cudaEventCreate(&g.ev[k]));
  for (int k = 0; k < 2; ++k) if (!g.gev[k]) CK(cudaEventCreate(&g.gev[k]));
  g.ready = true;
  return CPG_B200_OK;
}

int CPG_B200_FN(cpg_b200_free)(void) {        // releases the context of EVERY device
  Ctx* keep = cur_ctx;
  for (int d = 0; d < MAX_DEVICES; ++d) {
    cur_ctx = &ctxs[d];
    if (g.device < 0) continue;
    void* ptrs[] = {g.d_blob, g.d_cblob, g.d_gblob, g.d_gS0, g.d_mblob, g.d_dblob, g.d_mat_scratch, g.g_soly, g.g_dprim, g.g_dparams, g.g_dq, g.g_dl, g.g_du, g.d_tail_blob, g.d_counter, g.d_tail_count, g.d_tail_ids, g.d_tail_state, g.d_params, g.d_x0, g.d_y0, g.d_ws,
                    g.d_prim, g.d_dual, g.d_solx, g.d_soly, g.d_obj, g.d_pri, g.d_dua, g.d_iter, g.d_status};
    cudaSetDevice(g.device);
    for (void* p : ptrs) if (p) cudaFree(p);
    for (cudaEvent_t e : g.ev) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : g.gev) if (e) cudaEventDestroy(e);
    g = Ctx();
  }
  cur_ctx = keep;
  return CPG_B200_OK;
}

int CPG_B200_FN(cpg_solve_batch_device)(int B, const double* params, const double* x0, const double* y0,
                                        double* prim, double* dual, double* sol_x, double* sol_y,
                                        double* obj_val, int* iter, int* status, double* pri_res, double* dua_res,
                                        const CpgB200Settings* settings, void* stream_) {
  USE_DEVICE();
  if (B < 0 || !obj_val || !iter || !status || !pri_res || !dua_res) return CPG_B200_ERR_BAD_ARG;
  g.launches = 0;
  if (B == 0) return CPG_B200_OK;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CpgB200Settings s;
  if (settings) s = *settings; else CPG_B200_FN(cpg_b200_default_settings)(&s);
  int rc = ensure_tail(B);
  if (rc) return rc;
  cpgb200::Settings st;
  st.max_iter = s.max_iter; st.check_termination = s.check_termination; st.scaled_termination = s.scaled_termination;
  st.warm_start = s.warm_start && x0 && y0; st.adaptive_rho = s.adaptive_rho;
  st.adaptive_rho_interval = s.adaptive_rho_interval;
  if (st.adaptive_rho && !st.adaptive_rho_interval)        // osqp.c:267-279 (no profiling timer)
    st.adaptive_rho_interval = s.check_termination ? 4 * s.check_termination : 100;
  st.scaling = CPG_FAM_SCALING; st.pad = 0;
  st.eps_abs = s.eps_abs; st.eps_rel = s.eps_rel; st.eps_prim_inf = s.eps_prim_inf; st.eps_dual_inf = s.eps_dual_inf;
  st.alpha = s.alpha; st.adaptive_rho_tolerance = s.adaptive_rho_tolerance;
  cpgb200::BatchIO io;
  io.params = params; io.x0 = x0; io.y0 = y0; io.prim = prim; io.dual = dual; io.sol_x = sol_x; io.sol_y = sol_y;
  io.obj_val = obj_val; io.iter = iter; io.status = status; io.pri_res = pri_res; io.dua_res = dua_res;
  io.work_counter = g.d_counter; io.tail_count = g.d_tail_count; io.tail_ids = g.d_tail_ids;
  io.tail_state = g.d_tail_state; io.B = B; io.tail_capacity = g.tail_cap;
  io.ws = nullptr;
  if (st.warm_start && Fam::M > 0) {        // start points (z0 = A x0, y0 / rho) of the warm-started instances, main kernel only
    if (B > g.cap_ws) {
      g.cap_ws = 0;
      CK(cudaStreamSynchronize(stream));
      int rc = grow(&g.d_ws, (size_t)B * 2 * Fam::M);
      if (rc) return rc;
      g.cap_ws = B;
    }
    io.ws = g.d_ws;
  }
  CK(cudaMemsetAsync(g.d_counter, 0, sizeof(unsigned int), stream));
  CK(cudaMemsetAsync(g.d_tail_count, 0, sizeof(int), stream));
#if CPG_FAM_MATPAR
  // a batched parameter enters P or A: every instance is equilibrated, assembled and factored by its own warp (row f2)
  {
    int grid = g.n_sm;
    const int need = (B + Fam::MAT_WARPS - 1) / Fam::MAT_WARPS;
    if (grid > need) grid = need;
    CK(cudaEventRecord(g.ev[0], stream));
    cpgb200::admm_matpar_kernel<Fam><<<grid, Fam::MAT_WARPS * 32, MAT_SMEM_BYTES, stream>>>(g.d_cblob, g.d_tail_blob, g.d_mblob, g.d_mat_scratch, io, st);
    g.launches += 1;
    CK(cudaGetLastError());
    CK(cudaEventRecord(g.ev[1], stream));
    CK(cudaEventRecord(g.ev[2], stream));
    g.ev_solve = true;
    return CPG_B200_OK;
  }
#else        // (the shared-matrix kernels below are not compiled into a matrix-parameter library)
  // one persistent CTA per SM; a batch smaller than one wave of slots is still spread over all SMs (every warp pulls its
  // instances from the global counter), so that few warps share an SM's shared-memory bandwidth: lower latency
  int grid = g.n_sm;
#if CPG_FAM_BIG
  // the tile schedule of this family is larger than shared memory: every instance is solved on its own numeric factor by the
  // per-instance-factor kernel (cold start; x0 / y0 are not used on this path), tables read through L2
  {
    const CpgBlobHeader* Hh = reinterpret_cast<const CpgBlobHeader*>(CPG_B200_FN(cpg_blob_words));
    CK(cudaEventRecord(g.ev[0], stream));
    cpgb200::queue_all_kernel<<<(B + 255) / 256, 256, 0, stream>>>(io, TAIL_WORDS, Fam::N + 2 * Fam::M, Hh->rho);
    (void)grid;
  }
#elif CPG_FAM_DMMA
  // tensor-core main kernel: a group of four warps takes eight instances at a time (KKT factor shared by the batch:
  // the solve is a dense contraction over instances, mma.sync.m8n8k4.f64)
  const int need = (B + 7) / 8;
  if (grid > need) grid = need;
  CK(cudaEventRecord(g.ev[0], stream));
  cpgb200::admm_dmma_kernel<Fam><<<grid, Fam::DM_GROUPS * 128, DMMA_SMEM_BYTES, stream>>>(g.d_cblob, g.d_dblob, io, st);
#else
  const int need = (B + Fam::NI - 1) / Fam::NI;
  if (grid > need) grid = need;
  CK(cudaEventRecord(g.ev[0], stream));
  cpgb200::admm_multi_kernel<Fam><<<grid, Fam::WARPS * 32, SMEM_BYTES, stream>>>(g.d_blob, io, st);
#endif
  g.launches += 1;
  CK(cudaGetLastError());
  CK(cudaEventRecord(g.ev[1], stream));
  // instances that changed rho (or a constraint type) continue with their own factor; the kernel exits at once
  // when the hand-off queue is empty (no host round trip to find out)
  cpgb200::admm_tail_kernel<Fam><<<g.n_sm, Fam::TAIL_WARPS * 32, TAIL_SMEM_BYTES, stream>>>(g.d_cblob, g.d_tail_blob, io, st);
  g.launches += 1;
  CK(cudaGetLastError());
  CK(cudaEventRecord(g.ev[2], stream));
  g.ev_solve = true;
  return CPG_B200_OK;
#endif
}

int CPG_B200_FN(cpg_solve_batch_host)(int B, const double* params, const double* x0, const double* y0,
                                      double* prim, double* dual, double* sol_x, double* sol_y,
                                      double* obj_val, int* iter, int* status, double* pri_res, double* dua_res,
                                      const CpgB200Settings* settings) {
  USE_DEVICE();
  if (B < 0 || !obj_val || !iter || !status || !pri_res || !dua_res) return CPG_B200_ERR_BAD_ARG;
  if (B == 0) { g.launches = 0; return CPG_B200_OK; }
  const CpgBlobHeader* H = reinterpret_cast<const CpgBlobHeader*>(CPG_B200_FN(cpg_blob_words));
  if (H->npb > 0 && !params) return CPG_B200_ERR_BAD_ARG;
  int rc = ensure_staging(B);
  if (rc) return rc;
  cudaStream_t st = 0;
  if (H->npb > 0) CK(cudaMemcpyAsync(g.d_params, params, sizeof(double) * (size_t)B * H->npb, cudaMemcpyHostToDevice, st));
  const bool warm = x0 && y0;
  if (warm) {
    CK(cudaMemcpyAsync(g.d_x0, x0, sizeof(double) * (size_t)B * Fam::N, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(g.d_y0, y0, sizeof(double) * (size_t)B * Fam::M, cudaMemcpyHostToDevice, st));
  }
  // result rows: a pinned host buffer is written by the kernels directly (the stores are posted PCIe writes, so the
  // transfer of instance i overlaps the solves still running); anything else goes through the staging buffers
  const bool zc = settings ? settings->host_zero_copy != 0 : true;
  double* k_prim = zc ? mapped_view(prim) : nullptr;
  double* k_dual = zc ? mapped_view(dual) : nullptr;
  double* k_solx = zc ? mapped_view(sol_x) : nullptr;
  double* k_soly = zc ? mapped_view(sol_y) : nullptr;
  rc = CPG_B200_FN(cpg_solve_batch_device)(B, g.d_params, warm ? g.d_x0 : nullptr, warm ? g.d_y0 : nullptr,
                                           prim ? (k_prim ? k_prim : g.d_prim) : nullptr,
                                           dual ? (k_dual ? k_dual : g.d_dual) : nullptr,
                                           sol_x ? (k_solx ? k_solx : g.d_solx) : nullptr,
                                           sol_y ? (k_soly ? k_soly : g.d_soly) : nullptr,
                                           g.d_obj, g.d_iter, g.d_status, g.d_pri, g.d_dua, settings, st);
  if (rc) return rc;
  if (prim && !k_prim) CK(cudaMemcpyAsync(prim, g.d_prim, sizeof(double) * (size_t)B * H->n_prim, cudaMemcpyDeviceToHost, st));
  if (dual && !k_dual) CK(cudaMemcpyAsync(dual, g.d_dual, sizeof(double) * (size_t)B * H->n_dual, cudaMemcpyDeviceToHost, st));
  if (sol_x && !k_solx) CK(cudaMemcpyAsync(sol_x, g.d_solx, sizeof(double) * (size_t)B * Fam::N, cudaMemcpyDeviceToHost, st));
  if (sol_y && !k_soly) CK(cudaMemcpyAsync(sol_y, g.d_soly, sizeof(double) * (size_t)B * Fam::M, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(obj_val, g.d_obj, sizeof(double) * (size_t)B, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(pri_res, g.d_pri, sizeof(double) * (size_t)B, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(dua_res, g.d_dua, sizeof(double) * (size_t)B, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(iter, g.d_iter, sizeof(int) * (size_t)B, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(status, g.d_status, sizeof(int) * (size_t)B, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return CPG_B200_OK;
}

int CPG_B200_FN(cpg_solve_batch_host_multi)(int n_dev, const int* devices, int B, const double* params, const double* x0,
                                            const double* y0, double* prim, double* dual, double* sol_x, double* sol_y,
                                            double* obj_val, int* iter, int* status, double* pri_res, double* dua_res,
                                            const CpgB200Settings* settings) {
  if (n_dev <= 0 || n_dev > MAX_DEVICES || B < 0 || !obj_val || !iter || !status || !pri_res || !dua_res) return CPG_B200_ERR_BAD_ARG;
  const CpgBlobHeader* H = reinterpret_cast<const CpgBlobHeader*>(CPG_B200_FN(cpg_blob_words));
  std::vector<int> dev(n_dev), rc(n_dev, CPG_B200_OK);
  for (int k = 0; k < n_dev; ++k) {
    dev[k] = devices ? devices[k] : k;
    if (dev[k] < 0 || dev[k] >= MAX_DEVICES) return CPG_B200_ERR_BAD_ARG;
    for (int j = 0; j < k; ++j) if (dev[j] == dev[k]) return CPG_B200_ERR_BAD_ARG;     // one host thread per context
  }
  Ctx* caller = cur_ctx;
  std::vector<std::thread> th;
  for (int k = 0; k < n_dev; ++k) {
    th.emplace_back([&, k] {
      // contiguous shard [lo, hi) of the batch: instances never interact, so nothing is exchanged between the devices
      const long long lo = (long long)B * k / n_dev, hi = (long long)B * (k + 1) / n_dev;
      // select this thread's context; the first use of a device uploads the constants (later calls must not: cpg_b200_init queries
      // the device properties and re-uploads every table -- ~2 ms per device, and it would undo a cpg_b200_load_constants_all)
      int r = ctxs[dev[k]].ready ? CPG_B200_FN(cpg_b200_use_device)(dev[k]) : CPG_B200_FN(cpg_b200_init)(dev[k]);
      if (r == CPG_B200_OK && hi > lo) {
        auto at = [&](auto* p, size_t w) { return p ? p + (size_t)lo * w : p; };
        r = CPG_B200_FN(cpg_solve_batch_host)((int)(hi - lo), at(params, (size_t)H->npb), at(x0, (size_t)Fam::N), at(y0, (size_t)Fam::M),
                                              at(prim, (size_t)H->n_prim), at(dual, (size_t)H->n_dual), at(sol_x, (size_t)Fam::N),
                                              at(sol_y, (size_t)Fam::M), at(obj_val, 1), at(iter, 1), at(status, 1), at(pri_res, 1),
                                              at(dua_res, 1), settings);
      }
      rc[k] = r;
      if (r != CPG_B200_OK) snprintf(caller->err, sizeof(caller->err), "device %d: %.200s", dev[k], g.err);
    });
  }
  for (auto& t : th) t.join();
  for (int k = 0; k < n_dev; ++k) if (rc[k] != CPG_B200_OK) return rc[k];
  return CPG_B200_OK;
}

int CPG_B200_FN(cpg_gradient_batch_device)(int B, const double* sol_x, const double* sol_y, const double* dprim,
                                           double* dparams, double* dq, double* dl, double* du, void* stream_) {
  (void)sol_x;   // only enters dP / dA (matrix parameters are shared in this build)
  USE_DEVICE();
#if CPG_FAM_MATPAR
  snprintf(g.err, sizeof(g.err), "this family has per-instance matrix parameters: call cpg_gradient_batch_*_mat (it needs the parameter rows)");
  return CPG_B200_ERR_BAD_ARG;
#endif
#if !CPG_FAM_GRAD
  snprintf(g.err, sizeof(g.err), "the backward kernel is not generated for a family of this size (its factor workspace exceeds shared memory)");
  return CPG_B200_ERR_BAD_ARG;
#endif
  if (B < 0 || !sol_y || !dprim) return CPG_B200_ERR_BAD_ARG;
  g.launches = 0;
  if (B == 0) return CPG_B200_OK;
  cpgb200::GradIO io;
  io.sol_y = sol_y; io.dprim = dprim; io.dparams = dparams; io.dq = dq; io.dl = dl; io.du = du; io.S0 = g.d_gS0; io.B = B;
  int grid = g.n_sm;
  const int need = (B + Fam::GRAD_WARPS - 1) / Fam::GRAD_WARPS;
  if (grid > need) grid = need;
  CK(cudaEventRecord(g.gev[0], reinterpret_cast<cudaStream_t>(stream_)));
  cpgb200::qp_grad_kernel<Fam><<<grid, Fam::GRAD_WARPS * 32, GRAD_SMEM_BYTES, reinterpret_cast<cudaStream_t>(stream_)>>>(
      g.d_gblob, g.d_tail_blob, io);
  g.launches += 1;
  CK(cudaGetLastError());
  CK(cudaEventRecord(g.gev[1], reinterpret_cast<cudaStream_t>(stream_)));
  g.ev_grad = true;
  return CPG_B200_OK;
}

int CPG_B200_FN(cpg_gradient_batch_device_mat)(int B, const double* params, const double* sol_x, const double* sol_y,
                                               const double* dprim, double* dparams, double* dq, double* dl, double* du,
                                               double* dP, double* dA, void* stream_) {
  USE_DEVICE();
#if CPG_FAM_MATPAR
  if (B < 0 || !params || !sol_x || !sol_y || !dprim) return CPG_B200_ERR_BAD_ARG;
  g.launches = 0;
  if (B == 0) return CPG_B200_OK;
  cpgb200::GradIO io;
  io.sol_y = sol_y; io.dprim = dprim; io.dparams = dparams; io.dq = dq; io.dl = dl; io.du = du; io.S0 = g.d_gS0; io.B = B;
  io.params = params; io.sol_x = sol_x; io.dP = dP; io.dA = dA; io.mblob = g.d_mblob; io.a_scratch = g.d_mat_scratch;
  int grid = g.n_sm;
  const int need = (B + Fam::GRAD_WARPS - 1) / Fam::GRAD_WARPS;
  if (grid > need) grid = need;
  CK(cudaEventRecord(g.gev[0], reinterpret_cast<cudaStream_t>(stream_)));
  cpgb200::qp_grad_kernel<Fam, true><<<grid, Fam::GRAD_WARPS * 32, GRAD_SMEM_BYTES, reinterpret_cast<cudaStream_t>(stream_)>>>(
      g.d_gblob, g.d_tail_blob, io);
  g.launches += 1;
  CK(cudaGetLastError());
  CK(cudaEventRecord(g.gev[1], reinterpret_cast<cudaStream_t>(stream_)));
  g.ev_grad = true;
  return CPG_B200_OK;
#else
  (void)B; (void)params; (void)sol_x; (void)sol_y; (void)dprim; (void)dparams; (void)dq; (void)dl; (void)du; (void)dP; (void)dA; (void)stream_;
  snprintf(g.err, sizeof(g.err), "this library was generated without per-instance matrix parameters: call cpg_gradient_batch_*");
  return CPG_B200_ERR_BAD_ARG;
#endif
}

int CPG_B200_FN(cpg_gradient_batch_host_mat)(int B, const double* params, const double* sol_x, const double* sol_y,
                                             const double* dprim, double* dparams, double* dq, double* dl, double* du,
                                             double* dP, double* dA) {
  USE_DEVICE();
#if CPG_FAM_MATPAR
  if (B < 0 || !params || !sol_x || !sol_y || !dprim) return CPG_B200_ERR_BAD_ARG;
  if (B == 0) { g.launches = 0; return CPG_B200_OK; }
  const CpgBlobHeader* H = reinterpret_cast<const CpgBlobHeader*>(CPG_B200_FN(cpg_blob_words));
  const CpgMatHeader* MH = reinterpret_cast<const CpgMatHeader*>(CPG_B200_FN(cpg_mblob_words));
  const size_t szs[] = {(size_t)H->npb, (size_t)Fam::N, (size_t)Fam::M, (size_t)H->n_prim,          // in: params x y dprim
                        (size_t)H->npb, (size_t)Fam::N, (size_t)Fam::M, (size_t)Fam::M, (size_t)MH->nnzP, (size_t)MH->nnzA};
  const double* ins[] = {params, sol_x, sol_y, dprim};
  double* outs[] = {dparams, dq, dl, du, dP, dA};
  double* dev[10] = {nullptr};
  cudaStream_t st = 0;
  int rc = CPG_B200_OK;
  for (int k = 0; k < 10 && rc == CPG_B200_OK; ++k) {
    if (k >= 4 && !outs[k - 4]) continue;
    if (cudaMalloc(&dev[k], sizeof(double) * (size_t)B * (szs[k] ? szs[k] : 1)) != cudaSuccess) rc = CPG_B200_ERR_CUDA;
  }
  for (int k = 0; k < 4 && rc == CPG_B200_OK; ++k)
    if (cudaMemcpyAsync(dev[k], ins[k], sizeof(double) * (size_t)B * szs[k], cudaMemcpyHostToDevice, st) != cudaSuccess) rc = CPG_B200_ERR_CUDA;
  if (rc == CPG_B200_OK)
    rc = CPG_B200_FN(cpg_gradient_batch_device_mat)(B, dev[0], dev[1], dev[2], dev[3], dev[4], dev[5], dev[6], dev[7], dev[8], dev[9], st);
  for (int k = 4; k < 10 && rc == CPG_B200_OK; ++k)
    if (outs[k - 4] && cudaMemcpyAsync(outs[k - 4], dev[k], sizeof(double) * (size_t)B * szs[k], cudaMemcpyDeviceToHost, st) != cudaSuccess)
      rc = CPG_B200_ERR_CUDA;
  if (cudaStreamSynchronize(st) != cudaSuccess && rc == CPG_B200_OK) rc = CPG_B200_ERR_CUDA;
  if (rc == CPG_B200_ERR_CUDA && !g.err[0]) snprintf(g.err, sizeof(g.err), "CUDA error in cpg_gradient_batch_host_mat: %s", cudaGetErrorString(cudaGetLastError()));
  for (int k = 0; k < 10; ++k) if (dev[k]) cudaFree(dev[k]);
  return rc;
#else
  (void)B; (void)params; (void)sol_x; (void)sol_y; (void)dprim; (void)dparams; (void)dq; (void)dl; (void)du; (void)dP; (void)dA;
  snprintf(g.err, sizeof(g.err), "this library was generated without per-instance matrix parameters: call cpg_gradient_batch_*");
  return CPG_B200_ERR_BAD_ARG;
#endif
}

int CPG_B200_FN(cpg_gradient_batch_host)(int B, const double* sol_x, const double* sol_y, const double* dprim,
                                         double* dparams, double* dq, double* dl, double* du) {
  USE_DEVICE();
  if (B < 0 || !sol_y || !dprim) return CPG_B200_ERR_BAD_ARG;
  if (B == 0) { g.launches = 0; return CPG_B200_OK; }
  const CpgBlobHeader* H = reinterpret_cast<const CpgBlobHeader*>(CPG_B200_FN(cpg_blob_words));
  int rc;
  if (B > g.cap_G) {
    g.cap_G = 0;
    if ((rc = grow(&g.g_soly, (size_t)B * (Fam::M > 0 ? Fam::M : 1)))) return rc;
    if ((rc = grow(&g.g_dprim, (size_t)B * (H->n_prim > 0 ? H->n_prim : 1)))) return rc;
    if ((rc = grow(&g.g_dparams, (size_t)B * (H->npb > 0 ? H->npb : 1)))) return rc;
    if ((rc = grow(&g.g_dq, (size_t)B * Fam::N))) return rc;
    if ((rc = grow(&g.g_dl, (size_t)B * (Fam::M > 0 ? Fam::M : 1)))) return rc;
    if ((rc = grow(&g.g_du, (size_t)B * (Fam::M > 0 ? Fam::M : 1)))) return rc;
    g.cap_G = B;
  }
  cudaStream_t st = 0;
  CK(cudaMemcpyAsync(g.g_soly, sol_y, sizeof(double) * (size_t)B * Fam::M, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(g.g_dprim, dprim, sizeof(double) * (size_t)B * H->n_prim, cudaMemcpyHostToDevice, st));
  rc = CPG_B200_FN(cpg_gradient_batch_device)(B, nullptr, g.g_soly, g.g_dprim, dparams ? g.g_dparams : nullptr,
                                              dq ? g.g_dq : nullptr, dl ? g.g_dl : nullptr, du ? g.g_du : nullptr, st);
  if (rc) return rc;
  if (dparams) CK(cudaMemcpyAsync(dparams, g.g_dparams, sizeof(double) * (size_t)B * H->npb, cudaMemcpyDeviceToHost, st));
  if (dq) CK(cudaMemcpyAsync(dq, g.g_dq, sizeof(double) * (size_t)B * Fam::N, cudaMemcpyDeviceToHost, st));
  if (dl) CK(cudaMemcpyAsync(dl, g.g_dl, sizeof(double) * (size_t)B * Fam::M, cudaMemcpyDeviceToHost, st));
  if (du) CK(cudaMemcpyAsync(du, g.g_du, sizeof(double) * (size_t)B * Fam::M, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return CPG_B200_OK;
}

}  // extern "C"
