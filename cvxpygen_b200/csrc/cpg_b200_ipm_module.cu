// cpg_b200_ipm_module.cu -- C-ABI runtime of one generated IPM-CUDA solver library (SOCP families).
// Compiled once per problem family together with the generated cpg_ipm_family.h (compile-time sizes, table offsets) and
// cpg_ipm_blob.c (the constant tables).  Implements include/cpg_b200_socp.h; see that header for the reference
// interfaces it stands beside.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <thread>
#include <vector>

#include "cpg_ipm_family.h"
#include "cpg_b200_socp.h"
#include "ipm_kernel.cuh"

extern "C" const unsigned long long CPG_B200_FN(cpg_ipm_sblob_words)[];
extern "C" const unsigned int CPG_B200_FN(cpg_ipm_sblob_nbytes);
extern "C" const unsigned long long CPG_B200_FN(cpg_ipm_gblob_words)[];
extern "C" const unsigned int CPG_B200_FN(cpg_ipm_gblob_nbytes);

namespace {

using cpgipm::IpmIO;
using cpgipm::IpmSettings;
static_assert(sizeof(IpmSettings) == sizeof(CpgB200SocpSettings), "settings struct of the kernel and of the C ABI must match");
static_assert(cpgipm::SMEM_BYTES <= 232448 - 64,   /* 227 KB opt-in limit per CTA minus the kernel's 16 bytes of static shared memory */ "per-instance state + tables exceed the shared memory of one SM");

struct Ctx {
  bool ready = false;
  int device = -1, n_sm = 0;
  uint8_t *d_sblob = nullptr, *d_gblob = nullptr;
  double* d_best = nullptr;
  int* d_counter = nullptr;
  int cap_B = 0;
  double *d_params = nullptr, *d_prim = nullptr, *d_dual = nullptr, *d_x = nullptr, *d_y = nullptr, *d_z = nullptr, *d_s = nullptr;
  double *d_obj = nullptr, *d_pri = nullptr, *d_dua = nullptr;
  int *d_iter = nullptr, *d_status = nullptr;
  int launches = 0;
  cudaEvent_t ev[2] = {nullptr, nullptr};     // around the last ipm_kernel launch
  bool ev_solve = false;
  char err[256] = {0};
};
// one context per DEVICE; a host thread works on the context it selected last (cpg_b200_init / cpg_b200_use_device)
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

// one context per library and process, bound to ONE device (see include/cpg_b200_socp.h)
#define USE_DEVICE()                                                                              \
  do {                                                                                            \
    if (!g.ready) { snprintf(g.err, sizeof(g.err), "cpg_b200_init has not been called"); return CPG_B200_ERR_NOT_INIT; } \
    CK(cudaSetDevice(g.device));                                                                  \
  } while (0)

template <class T_>
int grow(T_** p, size_t count) {
  if (*p) cudaFree(*p);
  *p = nullptr;
  CK(cudaMalloc(p, sizeof(T_) * (count ? count : 1)));
  return CPG_B200_OK;
}

int ensure_staging(int B) {
  if (B <= g.cap_B) return CPG_B200_OK;
  int rc;
  g.cap_B = 0;                 // a failed allocation below must not leave a stale capacity behind
  if ((rc = grow(&g.d_params, (size_t)B * cpgipm::NPB))) return rc;
  if ((rc = grow(&g.d_prim, (size_t)B * cpgipm::NPRIM))) return rc;
  if ((rc = grow(&g.d_dual, (size_t)B * cpgipm::NDUAL))) return rc;
  if ((rc = grow(&g.d_x, (size_t)B * cpgipm::N))) return rc;
  if ((rc = grow(&g.d_y, (size_t)B * cpgipm::P))) return rc;
  if ((rc = grow(&g.d_z, (size_t)B * cpgipm::M))) return rc;
  if ((rc = grow(&g.d_s, (size_t)B * cpgipm::M))) return rc;
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

int CPG_B200_FN(cpg_b200_use_device)(int device) {
  if (device < 0 || device >= MAX_DEVICES) return CPG_B200_ERR_BAD_ARG;
  if (!ctxs[device].ready) { snprintf(g.err, sizeof(g.err), "cpg_b200_init(%d) has not been called", device); return CPG_B200_ERR_NOT_INIT; }
  cur_ctx = &ctxs[device];
  return CPG_B200_OK;
}

int CPG_B200_FN(cpg_b200_init)(int device) {
  if (device < 0 || device >= MAX_DEVICES) { snprintf(g.err, sizeof(g.err), "device index %d outside [0, %d)", device, MAX_DEVICES); return CPG_B200_ERR_BAD_ARG; }
  cur_ctx = &ctxs[device];              // this thread now works on the context of `device`
  if (g.ready) return CPG_B200_OK;
  CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  g.device = device; g.n_sm = prop.multiProcessorCount;
  const size_t sb = CPG_B200_FN(cpg_ipm_sblob_nbytes), gb = CPG_B200_FN(cpg_ipm_gblob_nbytes);
  CK(cudaMalloc(&g.d_sblob, sb));
  CK(cudaMalloc(&g.d_gblob, gb));
  CK(cudaMemcpy(g.d_sblob, CPG_B200_FN(cpg_ipm_sblob_words), sb, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(g.d_gblob, CPG_B200_FN(cpg_ipm_gblob_words), gb, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&g.d_best, sizeof(double) * (size_t)g.n_sm * cpgipm::BEST_STRIDE));
  CK(cudaMalloc(&g.d_counter, sizeof(int)));
  CK(cudaFuncSetAttribute(cpgipm::ipm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cpgipm::SMEM_BYTES));
  CK(cudaEventCreate(&g.ev[0])); CK(cudaEventCreate(&g.ev[1]));
  g.ready = true;
  return CPG_B200_OK;
}

int CPG_B200_FN(cpg_b200_free)(void) {        // releases the context of EVERY device
  Ctx* keep = cur_ctx;
  for (int d = 0; d < MAX_DEVICES; ++d) {
    cur_ctx = &ctxs[d];
    if (!g.ready) continue;
    void* ptrs[] = {g.d_sblob, g.d_gblob, g.d_best, g.d_counter, g.d_params, g.d_prim, g.d_dual, g.d_x, g.d_y, g.d_z, g.d_s,
                    g.d_obj, g.d_pri, g.d_dua, g.d_iter, g.d_status};
    cudaSetDevice(g.device);
    for (void* p : ptrs) if (p) cudaFree(p);
    for (cudaEvent_t e : g.ev) if (e) cudaEventDestroy(e);
    g = Ctx();
  }
  cur_ctx = keep;
  return CPG_B200_OK;
}

const char* CPG_B200_FN(cpg_b200_last_error)(void) { return g.err; }
int CPG_B200_FN(cpg_b200_launch_count)(void) { return g.launches; }

int CPG_B200_FN(cpg_socp_dims)(CpgB200SocpDims* out) {
  if (!out) return CPG_B200_ERR_BAD_ARG;
  out->n_var = cpgipm::N; out->n_eq = cpgipm::P; out->n_ineq = cpgipm::M; out->n_lp = cpgipm::L; out->n_soc = cpgipm::NSOC;
  out->n_param = cpgipm::NPB; out->n_prim = cpgipm::NPRIM; out->n_dual = cpgipm::NDUAL;
  out->threads_per_cta = cpgipm::T; out->smem_bytes = (int)cpgipm::SMEM_BYTES;
  return CPG_B200_OK;
}

int CPG_B200_FN(cpg_b200_kernel_times)(float* main_ms, float* tail_ms, float* grad_ms) {
  USE_DEVICE();
  if (main_ms) *main_ms = -1.f;
  if (tail_ms) *tail_ms = -1.f;
  if (grad_ms) *grad_ms = -1.f;
  if (g.ev_solve && main_ms) {
    CK(cudaEventSynchronize(g.ev[1]));
    CK(cudaEventElapsedTime(main_ms, g.ev[0], g.ev[1]));
  }
  return CPG_B200_OK;
}

int CPG_B200_FN(cpg_socp_load_constants)(const void* smem_blob, int smem_nbytes, const void* gmem_blob, int gmem_nbytes) {
  USE_DEVICE();
  if (!smem_blob || !gmem_blob || smem_nbytes != (int)CPG_B200_FN(cpg_ipm_sblob_nbytes) ||
      gmem_nbytes != (int)CPG_B200_FN(cpg_ipm_gblob_nbytes)) {
    snprintf(g.err, sizeof(g.err), "constant tables of a different layout: regenerate the code");
    return CPG_B200_ERR_BAD_ARG;
  }
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(g.d_sblob, smem_blob, smem_nbytes, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(g.d_gblob, gmem_blob, gmem_nbytes, cudaMemcpyHostToDevice));
  return CPG_B200_OK;
}

void CPG_B200_FN(cpg_socp_default_settings)(CpgB200SocpSettings* s) {
  if (!s) return;
  s->maxit = 100; s->pad_ = 0;
  s->feastol = 1e-8; s->abstol = 1e-8; s->reltol = 1e-8;
  s->feastol_inacc = 1e-4; s->abstol_inacc = 5e-5; s->reltol_inacc = 5e-5;
}

int CPG_B200_FN(cpg_socp_solve_batch_device)(int B, const double* params, double* prim, double* dual,
                                             double* sol_x, double* sol_y, double* sol_z, double* sol_s,
                                             double* obj_val, int* iter, int* status, double* pri_res, double* dua_res,
                                             const CpgB200SocpSettings* settings, void* stream) {
  USE_DEVICE();
  if (B < 0 || (B > 0 && (!prim || !dual || !obj_val || !iter || !status || !pri_res || !dua_res || (cpgipm::NPB > 0 && !params)))) {
    snprintf(g.err, sizeof(g.err), "null output pointer or negative batch size");
    return CPG_B200_ERR_BAD_ARG;
  }
  g.launches = 0;
  if (B == 0) return CPG_B200_OK;
  CpgB200SocpSettings st;
  if (settings) st = *settings; else CPG_B200_FN(cpg_socp_default_settings)(&st);
  IpmSettings ks;
  memcpy(&ks, &st, sizeof(ks));
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  CK(cudaMemsetAsync(g.d_counter, 0, sizeof(int), s));
  IpmIO io{B, params, prim, dual, sol_x, sol_y, sol_z, sol_s, obj_val, iter, status, pri_res, dua_res, g.d_best, g.d_counter};
  const int grid = B < g.n_sm ? B : g.n_sm;
  CK(cudaEventRecord(g.ev[0], s));
  cpgipm::ipm_kernel<<<grid, cpgipm::T, cpgipm::SMEM_BYTES, s>>>(g.d_sblob, g.d_gblob, ks, io);
  CK(cudaGetLastError());
  CK(cudaEventRecord(g.ev[1], s));
  g.ev_solve = true;
  g.launches = 1;
  return CPG_B200_OK;
}

int CPG_B200_FN(cpg_socp_solve_batch_host)(int B, const double* params, double* prim, double* dual,
                                           double* sol_x, double* sol_y, double* sol_z, double* sol_s,
                                           double* obj_val, int* iter, int* status, double* pri_res, double* dua_res,
                                           const CpgB200SocpSettings* settings) {
  USE_DEVICE();
  if (B < 0 || (B > 0 && (!prim || !dual || !obj_val || !iter || !status || !pri_res || !dua_res || (cpgipm::NPB > 0 && !params)))) {
    snprintf(g.err, sizeof(g.err), "null output pointer or negative batch size");
    return CPG_B200_ERR_BAD_ARG;
  }
  if (B == 0) { g.launches = 0; return CPG_B200_OK; }
  int rc;
  if ((rc = ensure_staging(B))) return rc;
  if (cpgipm::NPB > 0) CK(cudaMemcpy(g.d_params, params, sizeof(double) * (size_t)B * cpgipm::NPB, cudaMemcpyHostToDevice));
  rc = CPG_B200_FN(cpg_socp_solve_batch_device)(B, g.d_params, g.d_prim, g.d_dual, sol_x ? g.d_x : nullptr, sol_y ? g.d_y : nullptr,
                                                sol_z ? g.d_z : nullptr, sol_s ? g.d_s : nullptr, g.d_obj, g.d_iter, g.d_status,
                                                g.d_pri, g.d_dua, settings, nullptr);
  if (rc) return rc;
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(prim, g.d_prim, sizeof(double) * (size_t)B * cpgipm::NPRIM, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(dual, g.d_dual, sizeof(double) * (size_t)B * cpgipm::NDUAL, cudaMemcpyDeviceToHost));
  if (sol_x) CK(cudaMemcpy(sol_x, g.d_x, sizeof(double) * (size_t)B * cpgipm::N, cudaMemcpyDeviceToHost));
  if (sol_y) CK(cudaMemcpy(sol_y, g.d_y, sizeof(double) * (size_t)B * cpgipm::P, cudaMemcpyDeviceToHost));
  if (sol_z) CK(cudaMemcpy(sol_z, g.d_z, sizeof(double) * (size_t)B * cpgipm::M, cudaMemcpyDeviceToHost));
  if (sol_s) CK(cudaMemcpy(sol_s, g.d_s, sizeof(double) * (size_t)B * cpgipm::M, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(obj_val, g.d_obj, sizeof(double) * (size_t)B, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(iter, g.d_iter, sizeof(int) * (size_t)B, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(status, g.d_status, sizeof(int) * (size_t)B, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(pri_res, g.d_pri, sizeof(double) * (size_t)B, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(dua_res, g.d_dua, sizeof(double) * (size_t)B, cudaMemcpyDeviceToHost));
  return CPG_B200_OK;
}

int CPG_B200_FN(cpg_socp_solve_batch_host_multi)(int n_dev, const int* devices, int B, const double* params, double* prim, double* dual,
                                                 double* sol_x, double* sol_y, double* sol_z, double* sol_s, double* obj_val, int* iter,
                                                 int* status, double* pri_res, double* dua_res, const CpgB200SocpSettings* settings) {
  if (n_dev <= 0 || n_dev > MAX_DEVICES || B < 0) return CPG_B200_ERR_BAD_ARG;
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
      const long long lo = (long long)B * k / n_dev, hi = (long long)B * (k + 1) / n_dev;     // contiguous shard, nothing exchanged
      int r = ctxs[dev[k]].ready ? CPG_B200_FN(cpg_b200_use_device)(dev[k]) : CPG_B200_FN(cpg_b200_init)(dev[k]);     // (init only on first use)
      if (r == CPG_B200_OK && hi > lo) {
        auto at = [&](auto* p, size_t w) { return p ? p + (size_t)lo * w : p; };
        r = CPG_B200_FN(cpg_socp_solve_batch_host)((int)(hi - lo), at(params, (size_t)cpgipm::NPB), at(prim, (size_t)cpgipm::NPRIM),
                                                   at(dual, (size_t)cpgipm::NDUAL), at(sol_x, (size_t)cpgipm::N), at(sol_y, (size_t)cpgipm::P),
                                                   at(sol_z, (size_t)cpgipm::M), at(sol_s, (size_t)cpgipm::M), at(obj_val, 1), at(iter, 1),
                                                   at(status, 1), at(pri_res, 1), at(dua_res, 1), settings);
      }
      rc[k] = r;
      if (r != CPG_B200_OK) snprintf(caller->err, sizeof(caller->err), "device %d: %.200s", dev[k], g.err);
    });
  }
  for (auto& t : th) t.join();
  for (int k = 0; k < n_dev; ++k) if (rc[k] != CPG_B200_OK) return rc[k];
  return CPG_B200_OK;
}

}  // extern "C"
