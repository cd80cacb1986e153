/* oracle/osqp_ref_driver.c -- TEST INFRASTRUCTURE (oracle), not product code.
 *
 * Thin batch driver around the UNMODIFIED vendored OSQP 0.6.2 sources, which are
 * compiled where they lie under /root/reference (see oracle/Makefile); only the
 * resulting oracle/_ref/libosqp_ref.so travels.  It reproduces, per instance,
 * what cvxpygen's generated cpg_solve() does on the OSQP path
 * (cvxpygen/utils.py:1009-1052, cvxpygen/solvers/osqp.py:20-62):
 *     [osqp_update_P_A]  ->  osqp_update_lin_cost / osqp_update_bounds  ->  osqp_solve
 * (the matrix update first, like the reference's update table; before it the vectors are put back to their
 *  generation-time values so that scale_data inside osqp_update_P_A sees the pristine q whatever was solved before)
 * with the batch semantics "every instance is solved from the pristine
 * post-setup state" (rho reset to its initial value, cold start), because a
 * batch has no defined instance order.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.
 */
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <pthread.h>
#include "osqp.h"
#include "auxil.h"

typedef struct {
  int     max_iter;
  double  eps_abs, eps_rel, eps_prim_inf, eps_dual_inf;
  double  rho, sigma, alpha;
  int     scaling, adaptive_rho, adaptive_rho_interval;
  double  adaptive_rho_tolerance;
  int     scaled_termination, check_termination, warm_start, polish;
} RefSettings;

typedef struct {
  int n, m, nthreads;
  double rho0;
  double *q0, *l0, *u0;   /* generation-time vectors (pristine state for matrix updates) */
  OSQPWorkspace **work;   /* one workspace per thread */
  RefSettings s;
} RefOsqp;

void ref_osqp_default_settings(RefSettings *s) {
  /* cvxpygen's table, cvxpygen/solvers/osqp.py:102-115, on top of OSQP defaults
     (osqp_sources/include/constants.h:59-114) */
  OSQPSettings d;
  osqp_set_default_settings(&d);
  s->max_iter = 4000; s->eps_abs = 1e-3; s->eps_rel = 1e-3;
  s->eps_prim_inf = 1e-4; s->eps_dual_inf = 1e-4;
  s->rho = d.rho; s->sigma = d.sigma; s->alpha = d.alpha;
  s->scaling = (int)d.scaling; s->adaptive_rho = (int)d.adaptive_rho;
  s->adaptive_rho_interval = (int)d.adaptive_rho_interval;
  s->adaptive_rho_tolerance = d.adaptive_rho_tolerance;
  s->scaled_termination = 0; s->check_termination = 25;
  s->warm_start = 0; s->polish = 0;
}

static void to_osqp_settings(const RefSettings *s, OSQPSettings *o) {
  osqp_set_default_settings(o);
  o->max_iter = s->max_iter; o->eps_abs = s->eps_abs; o->eps_rel = s->eps_rel;
  o->eps_prim_inf = s->eps_prim_inf; o->eps_dual_inf = s->eps_dual_inf;
  o->rho = s->rho; o->sigma = s->sigma; o->alpha = s->alpha;
  o->scaling = s->scaling; o->adaptive_rho = s->adaptive_rho;
  o->adaptive_rho_interval = s->adaptive_rho_interval;
  o->adaptive_rho_tolerance = s->adaptive_rho_tolerance;
  o->scaled_termination = s->scaled_termination;
  o->check_termination = s->check_termination;
  o->warm_start = s->warm_start; o->polish = s->polish;
  o->verbose = 0;
}

RefOsqp *ref_osqp_setup(int n, int m,
                        const int *Pp, const int *Pi, const double *Px,
                        const double *q,
                        const int *Ap, const int *Ai, const double *Ax,
                        const double *l, const double *u,
                        const RefSettings *s, int nthreads) {
  RefOsqp *r = (RefOsqp *)calloc(1, sizeof(RefOsqp));
  int t;
  if (nthreads < 1) nthreads = 1;
  r->n = n; r->m = m; r->nthreads = nthreads; r->s = *s; r->rho0 = s->rho;
  r->work = (OSQPWorkspace **)calloc(nthreads, sizeof(OSQPWorkspace *));
  r->q0 = (double *)malloc(sizeof(double) * (n > 0 ? n : 1)); memcpy(r->q0, q, sizeof(double) * n);
  r->l0 = (double *)malloc(sizeof(double) * (m > 0 ? m : 1)); memcpy(r->l0, l, sizeof(double) * m);
  r->u0 = (double *)malloc(sizeof(double) * (m > 0 ? m : 1)); memcpy(r->u0, u, sizeof(double) * m);
  for (t = 0; t < nthreads; t++) {
    OSQPData data; OSQPSettings st;
    csc P, A;
    memset(&P, 0, sizeof(P)); memset(&A, 0, sizeof(A));
    P.m = n; P.n = n; P.p = (c_int *)Pp; P.i = (c_int *)Pi; P.x = (c_float *)Px;
    P.nzmax = Pp[n]; P.nz = -1;
    A.m = m; A.n = n; A.p = (c_int *)Ap; A.i = (c_int *)Ai; A.x = (c_float *)Ax;
    A.nzmax = Ap[n]; A.nz = -1;
    data.n = n; data.m = m; data.P = &P; data.A = &A;
    data.q = (c_float *)q; data.l = (c_float *)l; data.u = (c_float *)u;
    to_osqp_settings(s, &st);
    if (osqp_setup(&r->work[t], &data, &st) != 0) { free(r->work); free(r); return 0; }
  }
  return r;
}

void ref_osqp_free(RefOsqp *r) {
  int t;
  if (!r) return;
  for (t = 0; t < r->nthreads; t++) if (r->work[t]) osqp_cleanup(r->work[t]);
  free(r->q0); free(r->l0); free(r->u0);
  free(r->work); free(r);
}

/* scaling vectors etc. for cross-checking the product's offline pipeline */
void ref_osqp_get_scaling(RefOsqp *r, double *D, double *E, double *c) {
  OSQPWorkspace *w = r->work[0];
  memcpy(D, w->scaling->D, sizeof(double) * r->n);
  memcpy(E, w->scaling->E, sizeof(double) * r->m);
  *c = w->scaling->c;
}
int ref_osqp_adaptive_rho_interval(RefOsqp *r) { return (int)r->work[0]->settings->adaptive_rho_interval; }

/* Solve B instances.  q/l/u batches may be NULL (= keep the setup value).
 * x0/y0 non-NULL => warm start from them (osqp_warm_start semantics).
 * Threads (pthreads; libgomp is not in this image) pull chunks of 16 instances
 * from a shared counter, one OSQP workspace per thread.
 * Returns wall seconds spent in the solve loop. */
typedef struct {
  RefOsqp *r; int B; int tid;
  const double *qb, *lb, *ub, *x0, *y0;
  const double *Pb, *Ab; int nnzP, nnzA;     /* per-instance matrix values (CSC order) or NULL */
  double *x, *y, *obj; int *iter, *status; double *pri_res, *dua_res; int *rho_updates;
  int *next;
} Job;

static void solve_one(Job *j, OSQPWorkspace *w, int b) {
  RefOsqp *r = j->r; int n = r->n, m = r->m;
  if (w->settings->rho != r->rho0) osqp_update_rho(w, r->rho0);
  if (j->Pb || j->Ab) {
    osqp_update_lin_cost(w, r->q0);
    if (m > 0) osqp_update_bounds(w, r->l0, r->u0);
    /* 0.6.2 spells "the other matrix is NULL" (osqp_update_data_mat of OSQP v1) as osqp_update_P / osqp_update_A:
       each un-scales the data, overwrites its matrix, re-runs scale_data and re-factors (osqp.c:974-1156) */
    if (j->Pb && j->Ab) osqp_update_P_A(w, j->Pb + (size_t)b * j->nnzP, 0, j->nnzP, j->Ab + (size_t)b * j->nnzA, 0, j->nnzA);
    else if (j->Ab) osqp_update_A(w, j->Ab + (size_t)b * j->nnzA, 0, j->nnzA);
    else osqp_update_P(w, j->Pb + (size_t)b * j->nnzP, 0, j->nnzP);
  }
  if (j->qb) osqp_update_lin_cost(w, j->qb + (size_t)b * n);
  if (j->lb && j->ub) osqp_update_bounds(w, j->lb + (size_t)b * m, j->ub + (size_t)b * m);
  if (j->x0 && j->y0) {
    w->settings->warm_start = 1;
    osqp_warm_start(w, j->x0 + (size_t)b * n, j->y0 + (size_t)b * m);
  } else {
    w->settings->warm_start = 0;
  }
  osqp_solve(w);
  memcpy(j->x + (size_t)b * n, w->solution->x, sizeof(double) * n);
  memcpy(j->y + (size_t)b * m, w->solution->y, sizeof(double) * m);
  j->obj[b] = w->info->obj_val; j->iter[b] = (int)w->info->iter;
  j->status[b] = (int)w->info->status_val;
  j->pri_res[b] = w->info->pri_res; j->dua_res[b] = w->info->dua_res;
  if (j->rho_updates) j->rho_updates[b] = (int)w->info->rho_updates;
}

static void *worker(void *arg) {
  Job *j = (Job *)arg;
  OSQPWorkspace *w = j->r->work[j->tid];
  for (;;) {
    int b0 = __atomic_fetch_add(j->next, 16, __ATOMIC_RELAXED);
    int b1 = b0 + 16 < j->B ? b0 + 16 : j->B;
    if (b0 >= j->B) break;
    for (int b = b0; b < b1; b++) solve_one(j, w, b);
  }
  return 0;
}

static double run_jobs(RefOsqp *r, int B, Job *proto, int nthreads);

/* Same with per-instance matrices: Pb (B, nnzP) / Ab (B, nnzA) in CSC order, either may be NULL
 * (osqp_update_P_A with Px_new_idx = Ax_new_idx = NULL, osqp.c:1158). */
double ref_osqp_solve_batch_mat(RefOsqp *r, int B, const double *Pb, int nnzP, const double *Ab, int nnzA,
                                const double *qb, const double *lb, const double *ub,
                                double *x, double *y, double *obj,
                                int *iter, int *status, double *pri_res, double *dua_res,
                                int *rho_updates, int nthreads) {
  Job j; memset(&j, 0, sizeof(j));
  j.r = r; j.B = B; j.qb = qb; j.lb = lb; j.ub = ub; j.Pb = Pb; j.Ab = Ab; j.nnzP = nnzP; j.nnzA = nnzA;
  j.x = x; j.y = y; j.obj = obj; j.iter = iter; j.status = status;
  j.pri_res = pri_res; j.dua_res = dua_res; j.rho_updates = rho_updates;
  return run_jobs(r, B, &j, nthreads);
}

double ref_osqp_solve_batch(RefOsqp *r, int B,
                            const double *qb, const double *lb, const double *ub,
                            const double *x0, const double *y0,
                            double *x, double *y, double *obj,
                            int *iter, int *status, double *pri_res, double *dua_res,
                            int *rho_updates, int nthreads) {
  Job j; memset(&j, 0, sizeof(j));
  j.r = r; j.B = B; j.qb = qb; j.lb = lb; j.ub = ub; j.x0 = x0; j.y0 = y0;
  j.x = x; j.y = y; j.obj = obj; j.iter = iter; j.status = status;
  j.pri_res = pri_res; j.dua_res = dua_res; j.rho_updates = rho_updates;
  return run_jobs(r, B, &j, nthreads);
}

static double run_jobs(RefOsqp *r, int B, Job *proto, int nthreads) {
  struct timespec t0, t1;
  int next = 0, t;
  Job jobs[256]; pthread_t th[256];
  if (nthreads < 1) nthreads = 1;
  if (nthreads > r->nthreads) nthreads = r->nthreads;
  if (nthreads > 256) nthreads = 256;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (t = 0; t < nthreads; t++) { jobs[t] = *proto; jobs[t].tid = t; jobs[t].next = &next; }
  if (nthreads == 1) worker(&jobs[0]);
  else {
    for (t = 0; t < nthreads; t++) pthread_create(&th[t], 0, worker, &jobs[t]);
    for (t = 0; t < nthreads; t++) pthread_join(th[t], 0);
  }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}
