/* oracle/ecos_ref_driver.c -- TEST INFRASTRUCTURE (oracle), not product code.
 *
 * Batch driver around the UNMODIFIED vendored ECOS 2.0.8 sources (compiled where they lie under /root/reference,
 * see oracle/Makefile).  Per instance it does what cvxpygen's generated code does on the ECOS path
 * (cvxpygen/solvers/ecos.py:88-106): ECOS_setup once, then for every instance fresh un-equilibrated copies of
 * G, A, c, h, b into the arrays handed to setup (cpg_copy_all) + ECOS_updateData + ECOS_solve.
 * Settings = cvxpygen's table (cvxpygen/solvers/ecos.py:60-68): feastol = abstol = reltol = 1e-8, maxit = 100.
 */
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "ecos.h"

typedef struct {
  idxint n, m, p, l, ncones;
  idxint *q, *Gjc, *Gir, *Ajc, *Air;
  pfloat *Gpr, *Apr, *c, *h, *b;        /* working copies handed to ECOS (equilibrated in place) */
  pfloat *Gpr0, *Apr0, *c0, *h0, *b0;   /* pristine templates */
  idxint nnzG, nnzA;
  pwork* w;
} EcosRef;

static void* dup_(const void* src, size_t bytes) { void* d = malloc(bytes ? bytes : 1); if (bytes) memcpy(d, src, bytes); return d; }

EcosRef* ecos_ref_setup(long n, long m, long p, long l, long ncones, const long* q,
                        const double* Gpr, const long* Gjc, const long* Gir,
                        const double* Apr, const long* Ajc, const long* Air,
                        const double* c, const double* h, const double* b,
                        double feastol, double abstol, double reltol, long maxit) {
  EcosRef* r = (EcosRef*)calloc(1, sizeof(EcosRef));
  long i;
  r->n = n; r->m = m; r->p = p; r->l = l; r->ncones = ncones;
  r->q = (idxint*)malloc(sizeof(idxint) * (ncones ? ncones : 1));
  for (i = 0; i < ncones; i++) r->q[i] = q[i];
  r->nnzG = Gjc[n]; r->nnzA = p ? Ajc[n] : 0;
  r->Gjc = (idxint*)dup_(Gjc, sizeof(idxint) * (n + 1)); r->Gir = (idxint*)dup_(Gir, sizeof(idxint) * r->nnzG);
  r->Gpr0 = (pfloat*)dup_(Gpr, sizeof(pfloat) * r->nnzG); r->Gpr = (pfloat*)dup_(Gpr, sizeof(pfloat) * r->nnzG);
  if (p) {
    r->Ajc = (idxint*)dup_(Ajc, sizeof(idxint) * (n + 1)); r->Air = (idxint*)dup_(Air, sizeof(idxint) * r->nnzA);
    r->Apr0 = (pfloat*)dup_(Apr, sizeof(pfloat) * r->nnzA); r->Apr = (pfloat*)dup_(Apr, sizeof(pfloat) * r->nnzA);
    r->b0 = (pfloat*)dup_(b, sizeof(pfloat) * p); r->b = (pfloat*)dup_(b, sizeof(pfloat) * p);
  }
  r->c0 = (pfloat*)dup_(c, sizeof(pfloat) * n); r->c = (pfloat*)dup_(c, sizeof(pfloat) * n);
  r->h0 = (pfloat*)dup_(h, sizeof(pfloat) * m); r->h = (pfloat*)dup_(h, sizeof(pfloat) * m);
  r->w = ECOS_setup(n, m, p, l, ncones, r->q, 0, r->Gpr, r->Gjc, r->Gir, r->Apr, r->Ajc, r->Air, r->c, r->h, r->b);
  if (!r->w) { free(r); return 0; }
  r->w->stgs->feastol = feastol; r->w->stgs->abstol = abstol; r->w->stgs->reltol = reltol;
  r->w->stgs->maxit = maxit; r->w->stgs->verbose = 0;
  return r;
}

/* cb (B,n), hb (B,m), bb (B,p): per-instance vectors or NULL (= template).  Outputs x (B,n), y (B,p), z (B,m), s (B,m). */
double ecos_ref_solve_batch(EcosRef* r, long B, const double* cb, const double* hb, const double* bb,
                            double* x, double* y, double* z, double* s,
                            double* pcost, long* iter, long* exitflag, double* pres, double* dres) {
  struct timespec t0, t1;
  long k;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (k = 0; k < B; k++) {
    /* cpg_copy_all: fresh un-equilibrated data (ECOS preconditions in memory: cvxpygen/solvers/ecos.py:31) */
    memcpy(r->Gpr, r->Gpr0, sizeof(pfloat) * r->nnzG);
    if (r->p) memcpy(r->Apr, r->Apr0, sizeof(pfloat) * r->nnzA);
    memcpy(r->c, cb ? cb + k * r->n : r->c0, sizeof(pfloat) * r->n);
    memcpy(r->h, hb ? hb + k * r->m : r->h0, sizeof(pfloat) * r->m);
    if (r->p) memcpy(r->b, bb ? bb + k * r->p : r->b0, sizeof(pfloat) * r->p);
    ECOS_updateData(r->w, r->Gpr, r->Apr, r->c, r->h, r->b);
    exitflag[k] = ECOS_solve(r->w);
    memcpy(x + k * r->n, r->w->x, sizeof(pfloat) * r->n);
    if (r->p) memcpy(y + k * r->p, r->w->y, sizeof(pfloat) * r->p);
    memcpy(z + k * r->m, r->w->z, sizeof(pfloat) * r->m);
    memcpy(s + k * r->m, r->w->s, sizeof(pfloat) * r->m);
    pcost[k] = r->w->info->pcost; iter[k] = r->w->info->iter; pres[k] = r->w->info->pres; dres[k] = r->w->info->dres;
  }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}

/* The same with per-instance MATRIX values: Gb (B, nnzG), Ab (B, nnzA) in the CSC order of the templates, or NULL (= template).
 * This is the reference's path when a user parameter enters G or A: cpg_copy_all + ECOS_updateData on the raw values, which
 * re-equilibrates from scratch (cvxpygen/solvers/ecos.py:88-101, ecos/src/ecos.c:1648-1695, equil.c:210-342). */
double ecos_ref_solve_batch_mat(EcosRef* r, long B, const double* cb, const double* hb, const double* bb,
                                const double* Gb, const double* Ab,
                                double* x, double* y, double* z, double* s,
                                double* pcost, long* iter, long* exitflag, double* pres, double* dres) {
  struct timespec t0, t1;
  long k;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (k = 0; k < B; k++) {
    memcpy(r->Gpr, Gb ? Gb + k * r->nnzG : r->Gpr0, sizeof(pfloat) * r->nnzG);
    if (r->p) memcpy(r->Apr, Ab ? Ab + k * r->nnzA : r->Apr0, sizeof(pfloat) * r->nnzA);
    memcpy(r->c, cb ? cb + k * r->n : r->c0, sizeof(pfloat) * r->n);
    memcpy(r->h, hb ? hb + k * r->m : r->h0, sizeof(pfloat) * r->m);
    if (r->p) memcpy(r->b, bb ? bb + k * r->p : r->b0, sizeof(pfloat) * r->p);
    ECOS_updateData(r->w, r->Gpr, r->Apr, r->c, r->h, r->b);
    exitflag[k] = ECOS_solve(r->w);
    memcpy(x + k * r->n, r->w->x, sizeof(pfloat) * r->n);
    if (r->p) memcpy(y + k * r->p, r->w->y, sizeof(pfloat) * r->p);
    memcpy(z + k * r->m, r->w->z, sizeof(pfloat) * r->m);
    memcpy(s + k * r->m, r->w->s, sizeof(pfloat) * r->m);
    pcost[k] = r->w->info->pcost; iter[k] = r->w->info->iter; pres[k] = r->w->info->pres; dres[k] = r->w->info->dres;
  }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}

void ecos_ref_free(EcosRef* r) {
  if (!r) return;
  /* ECOS_cleanup frees only what setup allocated; the data arrays are ours */
  ECOS_cleanup(r->w, 0);
  free(r->q); free(r->Gjc); free(r->Gir); free(r->Gpr); free(r->Gpr0); free(r->c); free(r->c0); free(r->h); free(r->h0);
  if (r->p) { free(r->Ajc); free(r->Air); free(r->Apr); free(r->Apr0); free(r->b); free(r->b0); }
  free(r);
}
