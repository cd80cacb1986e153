/* cpg_b200.h -- C ABI of one generated ADMM-CUDA solver library (libcpg_b200_<name>.so).
 *
 * Drop-in boundary of the batched-solve hot path.  Plain pointers and sizes only -- no
 * torch / CUDA types in the signatures (`stream` is an opaque cudaStream_t handle, 0 = default).
 * Every function name is prefixed with the code-generation prefix exactly like the reference's
 * generated C (`<prefix>cpg_solve`, cvxpygen/utils.py:1087-1141); CPG_B200_PREFIX is empty by default.
 *
 * Part A keeps the reference's generated single-instance interface (what its pybind module
 * binds, cvxpygen/utils.py:1194-1270, 1331-1412):
 *     void <p>cpg_update_<param>(cpg_int idx, cpg_float val)      cvxpygen/utils.py:904-935
 *     void <p>cpg_solve()                                         cvxpygen/utils.py:1009-1052
 *     void <p>cpg_set_solver_default_settings()                   cvxpygen/utils.py:1069-1076
 *     void <p>cpg_set_solver_<setting>(value)                     cvxpygen/utils.py:1077-1084
 *     globals  <p>CPG_Prim, <p>CPG_Dual, <p>CPG_Info, <p>CPG_Result   cvxpygen/utils.py:745-798
 *   These are emitted per family into c/include/cpg_solve.h + c/include/cpg_workspace.h because
 *   their names depend on the user's parameter / variable names; they run a batch of one.
 *
 * Part B (this header) is the NEW batched entry the reference lacks (SURVEY section 8b, row "new").
 * All functions return 0 on success or a CPG_B200_ERR_* code; nothing aborts, nothing falls back to
 * a CPU path: without a CUDA device cpg_b200_init fails with CPG_B200_ERR_CUDA.
 */
#ifndef CPG_B200_H
#define CPG_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#ifndef CPG_B200_PREFIX
#define CPG_B200_PREFIX
#endif
#define CPG_B200_CAT_(a, b) a##b
#define CPG_B200_CAT(a, b) CPG_B200_CAT_(a, b)
#define CPG_B200_FN(name) CPG_B200_CAT(CPG_B200_PREFIX, name)

enum {
  CPG_B200_OK = 0,
  CPG_B200_ERR_CUDA = 1,        /* a CUDA runtime call failed: see cpg_b200_last_error() */
  CPG_B200_ERR_NOT_INIT = 2,    /* cpg_b200_init() has not been called */
  CPG_B200_ERR_BAD_ARG = 3,     /* null pointer / negative size */
  CPG_B200_ERR_TAIL_OVERFLOW = 4 /* more hand-offs to the refactorisation kernel than its queue holds */
};

/* OSQP status values reported per instance in `status` (osqp_sources/include/constants.h:18-30) */
enum {
  CPG_B200_SOLVED = 1, CPG_B200_SOLVED_INACCURATE = 2,
  CPG_B200_PRIMAL_INFEASIBLE_INACCURATE = 3, CPG_B200_DUAL_INFEASIBLE_INACCURATE = 4,
  CPG_B200_MAX_ITER_REACHED = -2, CPG_B200_PRIMAL_INFEASIBLE = -3, CPG_B200_DUAL_INFEASIBLE = -4,
  CPG_B200_NON_CVX = -7, CPG_B200_UNSOLVED = -10
};

/* Solver settings: the table cvxpygen exposes for OSQP (cvxpygen/solvers/osqp.py:102-115) plus the
 * OSQP defaults that shape the iteration (osqp_sources/include/constants.h:59-114).
 * rho and sigma are NOT here: they are baked into the KKT factor at generation time. */
typedef struct {
  int max_iter;               /* 4000 */
  int check_termination;      /* 25   */
  int scaled_termination;     /* 0    */
  int warm_start;             /* 0 for batches (cold start); 1 uses x0/y0 */
  int adaptive_rho;           /* 1    */
  int adaptive_rho_interval;  /* 0 = 4*check_termination, as OSQP without a timer (osqp.c:267-279) */
  int scaling;                /* read-only: number of Ruiz iterations used at generation time */
  int host_zero_copy;         /* 1: cpg_solve_batch_host lets the kernels store result rows straight into PINNED host
                                 buffers (posted PCIe writes overlapping the solves); 0 or pageable memory: staging + D2H */
  double eps_abs, eps_rel;            /* 1e-3 */
  double eps_prim_inf, eps_dual_inf;  /* 1e-4 */
  double alpha;                       /* 1.6  */
  double adaptive_rho_tolerance;      /* 5    */
} CpgB200Settings;

/* Problem-family dimensions baked into the library. */
typedef struct {
  int n_var, n_con;        /* canonical QP: x in R^n_var, l <= A x <= u in R^n_con            */
  int n_param;             /* doubles per instance in `params` (batched user parameters only)  */
  int n_prim, n_dual;      /* doubles per instance in `prim` / `dual`                          */
  int blob_bytes;          /* shared-memory constants blob                                     */
  int warps_per_cta, smem_bytes;
} CpgB200Dims;

/* One context per DEVICE (up to 16).  cpg_b200_init(device) creates / refreshes that device's context and makes it the
 * calling THREAD's current one; every other entry point works on the calling thread's current context and makes its device
 * current (cudaSetDevice) before it touches CUDA.  cpg_b200_use_device switches a thread to an initialised context.
 * A multi-GPU caller therefore runs one host thread per device on the same loaded library; cpg_solve_batch_host_multi is
 * that loop (contiguous shards of the batch, no data exchanged between devices). */
int  CPG_B200_FN(cpg_b200_init)(int device);                 /* upload constants, allocate queues  */
int  CPG_B200_FN(cpg_b200_use_device)(int device);
int  CPG_B200_FN(cpg_b200_free)(void);
int  CPG_B200_FN(cpg_b200_dims)(CpgB200Dims* out);
void CPG_B200_FN(cpg_b200_default_settings)(CpgB200Settings* s);
const char* CPG_B200_FN(cpg_b200_last_error)(void);
int  CPG_B200_FN(cpg_b200_launch_count)(void);               /* kernels launched by the last solve call */
/* Device time of the kernels of the LAST solve / gradient call, from CUDA events recorded on the caller's stream around
 * each launch (blocks until they have completed): main solve kernel (admm_multi_kernel or admm_matpar_kernel), the
 * refactorisation kernel (admm_tail_kernel), the backward kernel (qp_grad_kernel); -1 where nothing was launched.
 * This is what bench.py's roofline block divides the algorithmic bytes by. */
int  CPG_B200_FN(cpg_b200_kernel_times)(float* main_ms, float* tail_ms, float* grad_ms);
/* Replace the constants blob (shared parameters changed => host re-ran the offline setup). */
int  CPG_B200_FN(cpg_b200_load_constants)(const void* blob, int nbytes);
/* Replace EVERY constants table after a shared-parameter update (role of osqp_update_data_mat -> re-scale + refactor,
 * osqp_sources/src/osqp.c:1158-1264, done on the host by the offline pipeline): main blob, its compact copy for the
 * tail kernel, the re-factorisation tables, the backward-pass blob and its KKT slot values.  The sparsity structure
 * must be the one the library was generated for (same generated solve code); sizes are checked. */
int  CPG_B200_FN(cpg_b200_load_constants_all)(const void* blob, int nbytes, const void* cblob, int cnbytes,
                                              const void* tail_blob, int tnbytes, const void* gblob, int gnbytes,
                                              const void* gS0, int snbytes);

/* Libraries whose main kernel is the FP64 tensor-core variant (admm_dmma_kernel; chosen at generation time when the factor is
 * shared by the batch and its tables fit in shared memory): replace the tables of that solve after a shared-parameter update.
 * No-op (CPG_B200_OK) for the other libraries. */
int  CPG_B200_FN(cpg_b200_load_dmma_constants)(const void* dblob, int nbytes);

/* Families generated with per-instance MATRIX parameters (a batched parameter enters P or A): replace the
 * tables of the per-instance osqp_update_data_mat path (canonicalisation maps of the P / A entries, KKT slot maps, the
 * round-trip base values) after a SHARED parameter changed.  Error for libraries generated without such parameters.
 * In these families `params` rows carry the matrix parameters too, and every instance is re-equilibrated
 * (scale_data, osqp_sources/src/scaling.c:44-156), assembled and factored (kkt.c:184-212, qdldl.c:72-233) on the GPU. */
int  CPG_B200_FN(cpg_b200_load_mat_constants)(const void* mblob, int nbytes);

/* Batched solve, DEVICE buffers (row-major, one instance per row), asynchronous on `stream`.
 *   params (B, n_param) in        x0 (B, n_var) / y0 (B, n_con) optional warm start (NULL = cold)
 *   prim (B, n_prim), dual (B, n_dual) out; sol_x (B, n_var), sol_y (B, n_con) optional (NULL = skip)
 *   obj_val, pri_res, dua_res: (B) double;  iter, status: (B) int                               */
int CPG_B200_FN(cpg_solve_batch_device)(int B, const double* params, const double* x0, const double* y0,
                                        double* prim, double* dual, double* sol_x, double* sol_y,
                                        double* obj_val, int* iter, int* status,
                                        double* pri_res, double* dua_res,
                                        const CpgB200Settings* settings, void* stream);

/* Same with HOST buffers: H2D of params, solve, D2H of the results, synchronous.  This is what the
 * generated cpg_solve()/cpg_module.solve_batch call. */
int CPG_B200_FN(cpg_solve_batch_host)(int B, const double* params, const double* x0, const double* y0,
                                      double* prim, double* dual, double* sol_x, double* sol_y,
                                      double* obj_val, int* iter, int* status,
                                      double* pri_res, double* dua_res,
                                      const CpgB200Settings* settings);

/* The same on several devices of one node from ONE call: `devices` lists n_dev distinct device indices (NULL = 0 .. n_dev-1);
 * instance rows [B k / n_dev, B (k+1) / n_dev) go to devices[k], each from its own host thread (initialising that device's
 * context on first use).  No collective: instances never interact.  Returns the first non-zero shard code. */
int CPG_B200_FN(cpg_solve_batch_host_multi)(int n_dev, const int* devices, int B, const double* params, const double* x0,
                                            const double* y0, double* prim, double* dual, double* sol_x, double* sol_y,
                                            double* obj_val, int* iter, int* status, double* pri_res, double* dua_res,
                                            const CpgB200Settings* settings);

/* Batched backward pass (gradient=True): differentiates the QP solution map through its KKT system.
 * Reference counterpart, one instance at a time: <p>cpg_update_d<var>(idx, val) + <p>cpg_gradient()
 * (cvxpygen/writer.py:222-312) -> cpg_osqp_gradient() (templates/cpg_osqp_grad_compute.c.jinja2:432-531); the pybind
 * entry is cpg_module.gradient(vdelta, gsol, use_sol) (cvxpygen/utils.py:1272-1328).
 *   sol_x (B, n_var), sol_y (B, n_con): canonical solution returned by the forward solve (sol_x may be NULL: it only
 *                                       enters the gradients of matrix parameters, which are shared in this build)
 *   dprim   (B, n_prim): upstream gradient w.r.t. the user-level primal variables, same layout as `prim`
 *   dparams (B, n_param): OUT gradient w.r.t. the batched user parameters, same layout as `params`
 *   dq (B, n_var), dl, du (B, n_con): optional canonical gradients (NULL = skip)                               */
int CPG_B200_FN(cpg_gradient_batch_device)(int B, const double* sol_x, const double* sol_y, const double* dprim,
                                           double* dparams, double* dq, double* dl, double* du, void* stream);
int CPG_B200_FN(cpg_gradient_batch_host)(int B, const double* sol_x, const double* sol_y, const double* dprim,
                                         double* dparams, double* dq, double* dl, double* du);

/* The same for families generated with per-instance MATRIX parameters (SURVEY rows a16 + f2).  Reference counterpart:
 * the P / A branch of <p>cpg_gradient() (cvxpygen/writer.py:240-263: cpg_P_to_K, cpg_A_to_K, cpg_ldl_numeric) followed by
 * cpg_osqp_gradient() and the un-canonicalisation of dP / dA through canon_P_map / canon_A_map (writer.py:292-303).
 *   params (B, n_param): the parameter rows of the forward solve (each instance's P and A are canonicalised from them)
 *   sol_x required;  dP (B, nnz(P upper)), dA (B, nnz(A)): optional canonical matrix gradients in CSC order (NULL = skip)
 * Error for libraries generated without such parameters (and cpg_gradient_batch_* is an error for those with). */
int CPG_B200_FN(cpg_gradient_batch_device_mat)(int B, const double* params, const double* sol_x, const double* sol_y,
                                               const double* dprim, double* dparams, double* dq, double* dl, double* du,
                                               double* dP, double* dA, void* stream);
int CPG_B200_FN(cpg_gradient_batch_host_mat)(int B, const double* params, const double* sol_x, const double* sol_y,
                                             const double* dprim, double* dparams, double* dq, double* dl, double* du,
                                             double* dP, double* dA);

#ifdef __cplusplus
}
#endif
#endif /* CPG_B200_H */
