/* cpg_b200_socp.h -- C ABI of one generated IPM-CUDA solver library (SOCP families, libcpg_b200.so of an
 * `generate_code(..., solver='IPM-CUDA')` directory).
 *
 * Same conventions as cpg_b200.h: plain pointers and sizes, `stream` is an opaque cudaStream_t (0 = default), every
 * function name carries the code-generation prefix (cvxpygen/utils.py:1087-1141), every function returns 0 or a
 * CPG_B200_ERR_* code, nothing falls back to a CPU path.
 *
 * What it replaces, one instance at a time, in the reference's generated ECOS code (cvxpygen/solvers/ecos.py:88-117):
 *     cpg_copy_all + ECOS_updateData(ecos_workspace, G->x, A->x, c, h, b)      ecos/src/ecos.c:1648-1694
 *     ecos_flag = ECOS_solve(ecos_workspace)                                   ecos/src/ecos.c:1075-1607
 *     cpg_retrieve_prim / cpg_retrieve_dual / cpg_retrieve_info                cvxpygen/utils.py:950-985
 * and the settings table  feastol, abstol, reltol, feastol_inacc, abstol_inacc, reltol_inacc, maxit
 * (cvxpygen/solvers/ecos.py:60-68).
 *
 * A batched user parameter may enter c, b, h AND the matrices G, A: a library generated with such a parameter in `batch_params`
 * canonicalises every instance's G / A values from its parameter row and does what ECOS_updateData does with new values --
 * set_equilibration from scratch (ecos/src/equil.c:210-342) -- inside the kernel; nothing changes in this interface (the rows
 * of `params` are simply longer: matrix parameters in column-major order, like cvxpy's).
 */
#ifndef CPG_B200_SOCP_H
#define CPG_B200_SOCP_H

#ifdef __cplusplus
extern "C" {
#endif

#ifndef CPG_B200_PREFIX
#define CPG_B200_PREFIX
#endif
#ifndef CPG_B200_FN
#define CPG_B200_CAT_(a, b) a##b
#define CPG_B200_CAT(a, b) CPG_B200_CAT_(a, b)
#define CPG_B200_FN(name) CPG_B200_CAT(CPG_B200_PREFIX, name)
#endif

#ifndef CPG_B200_H
enum { CPG_B200_OK = 0, CPG_B200_ERR_CUDA = 1, CPG_B200_ERR_NOT_INIT = 2, CPG_B200_ERR_BAD_ARG = 3 };
#endif

/* ECOS exit flags reported per instance in `status` (ecos/include/ecos.h:87-95); +10 = "close to" (inaccurate) */
enum {
  CPG_B200_ECOS_OPTIMAL = 0, CPG_B200_ECOS_PINF = 1, CPG_B200_ECOS_DINF = 2, CPG_B200_ECOS_INACC_OFFSET = 10,
  CPG_B200_ECOS_MAXIT = -1, CPG_B200_ECOS_NUMERICS = -2, CPG_B200_ECOS_OUTCONE = -3, CPG_B200_ECOS_FATAL = -7
};

typedef struct {
  int maxit;                                            /* 100  */
  int pad_;
  double feastol, abstol, reltol;                       /* 1e-8 */
  double feastol_inacc, abstol_inacc, reltol_inacc;     /* 1e-4, 5e-5, 5e-5 */
} CpgB200SocpSettings;

typedef struct {
  int n_var, n_eq, n_ineq;      /* canonical SOCP: min c'x  s.t.  A x = b (n_eq rows),  h - G x in K (n_ineq rows) */
  int n_lp, n_soc;              /* K = R_+^{n_lp} x Q^{q_1} x ... x Q^{q_{n_soc}}                                   */
  int n_param;                  /* doubles per instance in `params` (batched user parameters only)                  */
  int n_prim, n_dual;           /* doubles per instance in `prim` / `dual`                                          */
  int threads_per_cta, smem_bytes;
} CpgB200SocpDims;

int  CPG_B200_FN(cpg_b200_init)(int device);                 /* upload the constant tables, allocate scratch */
int  CPG_B200_FN(cpg_b200_free)(void);
const char* CPG_B200_FN(cpg_b200_last_error)(void);
int  CPG_B200_FN(cpg_b200_launch_count)(void);               /* kernels launched by the last solve call */
/* Device time of the last ipm_kernel launch from CUDA events on the caller's stream (blocks until it has completed);
 * tail_ms / grad_ms are -1 (same signature as the QP libraries' entry).  One context per library and process, bound
 * One context per DEVICE; cpg_b200_init(device) makes that device's context the calling thread's current one, every entry
 * point works on the calling thread's context (cpg_b200_use_device switches). */
int  CPG_B200_FN(cpg_b200_kernel_times)(float* main_ms, float* tail_ms, float* grad_ms);
int  CPG_B200_FN(cpg_b200_use_device)(int device);
int  CPG_B200_FN(cpg_socp_dims)(CpgB200SocpDims* out);
/* Replace the constant tables after a change of SHARED (non-batched) user parameters: the host re-runs the offline setup
 * (equilibration, KKT base image, affine maps) and uploads both images.  Role of ECOS_updateData for data every instance
 * shares (cvxpygen/solvers/ecos.py:88-117, ecos/src/ecos.c:1648-1760).  The sizes must equal those compiled in: a change of
 * the sparsity structure needs regenerated code.  Synchronises the device. */
int  CPG_B200_FN(cpg_socp_load_constants)(const void* smem_blob, int smem_nbytes, const void* gmem_blob, int gmem_nbytes);
void CPG_B200_FN(cpg_socp_default_settings)(CpgB200SocpSettings* s);

/* Batched solve, DEVICE buffers (row-major, one instance per row), asynchronous on `stream`.
 *   params (B, n_param) in
 *   prim (B, n_prim), dual (B, n_dual) out: user-level variables and constraint duals
 *   sol_x (B, n_var), sol_y (B, n_eq), sol_z (B, n_ineq), sol_s (B, n_ineq): canonical solution, each optional (NULL)
 *   obj_val, pri_res, dua_res: (B) double;  iter, status: (B) int                                               */
int CPG_B200_FN(cpg_socp_solve_batch_device)(int B, const double* params, double* prim, double* dual,
                                             double* sol_x, double* sol_y, double* sol_z, double* sol_s,
                                             double* obj_val, int* iter, int* status, double* pri_res, double* dua_res,
                                             const CpgB200SocpSettings* settings, void* stream);
/* Same with HOST buffers: H2D of params, solve, D2H of the results, synchronous. */
int CPG_B200_FN(cpg_socp_solve_batch_host)(int B, const double* params, double* prim, double* dual,
                                           double* sol_x, double* sol_y, double* sol_z, double* sol_s,
                                           double* obj_val, int* iter, int* status, double* pri_res, double* dua_res,
                                           const CpgB200SocpSettings* settings);

/* The same on several devices of one node from ONE call: contiguous shards of the batch, one host thread per device inside the
 * library (devices = NULL: 0 .. n_dev-1; distinct indices); no collective.  Returns the first non-zero shard code. */
int CPG_B200_FN(cpg_socp_solve_batch_host_multi)(int n_dev, const int* devices, int B, const double* params, double* prim, double* dual,
                                                 double* sol_x, double* sol_y, double* sol_z, double* sol_s, double* obj_val, int* iter,
                                                 int* status, double* pri_res, double* dua_res, const CpgB200SocpSettings* settings);

#ifdef __cplusplus
}
#endif
#endif /* CPG_B200_SOCP_H */
