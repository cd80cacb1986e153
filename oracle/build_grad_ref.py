#!/usr/bin/env python
"""oracle/build_grad_ref.py -- TEST INFRASTRUCTURE (oracle), not product code.

Builds the REFERENCE's own generated-C backward pass for one problem family into
oracle/_ref/libgrad_ref_<family>.so.  Runs only where /root/reference exists (the build container):

  * the C sources are RENDERED from the reference's templates where they lie
    (cvxpygen/templates/cpg_osqp_grad_compute.{c,h}.jinja2, cpg_osqp_grad_workspace.h.jinja2) with the reference's own
    jinja environment and literal writers (cvxpygen/utils.py is imported by file path: it needs numpy, scipy and
    jinja2 only -- the cvxpygen package itself cannot be imported here because cvxpy is absent);
  * the static workspace follows cvxpygen/writer.py:354-416 (_write_gradient_workspace_def; writer.py imports cvxpy,
    so the entry list is restated below) and the first-call initialisation follows writer.py:232-251;
  * QDLDL is compiled from the vendored sources (osqp_sources/lin_sys/direct/qdldl/qdldl_sources/src/qdldl.c).
Nothing is copied into the repository: rendered files go to a temporary directory, only the .so is kept
(oracle/_ref/ is git-ignored but travels to the GPU box).

The driver processes instances SEQUENTIALLY through one workspace exactly as the reference does (the LDL' factor is
up/down-dated from one call's active set to the next, cpg_osqp_grad_compute.c.jinja2:437-454).
"""
import importlib.util
import os
import subprocess
import sys
import tempfile

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get('CPG_REFERENCE', '/root/reference')
QDLDL = os.path.join(REF, 'cvxpygen/solvers/osqp-python/osqp_sources/lin_sys/direct/qdldl/qdldl_sources')


def _ref_utils():
    spec = importlib.util.spec_from_file_location('cvxpygen_ref_utils', os.path.join(REF, 'cvxpygen', 'utils.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


DRIVER = r'''
#include <string.h>
#include "cpg_osqp_grad_workspace.h"
#include "cpg_osqp_grad_compute.h"
cpg_float sol_x[%(n)d];
cpg_float sol_y[%(m)d];
/* B instances, row-major; returns dq (B,n), dl (B,m), du (B,m).  First call initialises the factor
   (cvxpygen/writer.py:232-251). */
void grad_ref_batch(int B, const double* x, const double* y, const double* dx, double* dq, double* dl, double* du) {
  int b, i, j, k;
  for (b = 0; b < B; b++) {
    memcpy(sol_x, x + (size_t)b * %(n)d, sizeof(double) * %(n)d);
    memcpy(sol_y, y + (size_t)b * %(m)d, sizeof(double) * %(m)d);
    for (i = 0; i < %(n)d; i++) CPG_OSQP_Grad.dx[i] = dx[(size_t)b * %(n)d + i];
    if (CPG_OSQP_Grad.init) {
      cpg_ldl_symbolic();
      cpg_ldl_numeric();
      for (j = 0; j < %(N)d - 1; j++)
        for (k = CPG_OSQP_Grad.L->p[j]; k < CPG_OSQP_Grad.L->p[j + 1]; k++) {
          i = CPG_OSQP_Grad.L->i[k];
          CPG_OSQP_Grad.Lmask[(2 * %(N)d - 3 - j) * j / 2 + i - 1] = 1;
        }
      CPG_OSQP_Grad.init = 0;
    }
    cpg_osqp_gradient();
    memcpy(dq + (size_t)b * %(n)d, CPG_OSQP_Grad.dq, sizeof(double) * %(n)d);
    memcpy(dl + (size_t)b * %(m)d, CPG_OSQP_Grad.dl, sizeof(double) * %(m)d);
    memcpy(du + (size_t)b * %(m)d, CPG_OSQP_Grad.du, sizeof(double) * %(m)d);
  }
}
/* Per-instance matrices (Px (B, nnzP) upper triangle, Ax (B, nnzA), CSC order): what the generated cpg_gradient() does
   when P / A are outdated (cvxpygen/writer.py:240-263): cpg_P_to_K / cpg_A_to_K, cpg_ldl_numeric, a[] = 1, then
   cpg_osqp_gradient().  Also returns dP (B, nnzP) and dA (B, nnzA). */
void grad_ref_batch_mat(int B, const double* Px, const double* Ax, const double* x, const double* y, const double* dx,
                        double* dq, double* dl, double* du, double* dP, double* dA) {
  int b, i, j, k;
  const int nnzP = CPG_OSQP_Grad.dP->p[%(n)d], nnzA = CPG_OSQP_Grad.dA->p[%(n)d];
  for (b = 0; b < B; b++) {
    cpg_grad_csc Pm = *CPG_OSQP_Grad.dP, Am = *CPG_OSQP_Grad.dA;
    Pm.x = (cpg_grad_float*)(Px + (size_t)b * nnzP);
    Am.x = (cpg_grad_float*)(Ax + (size_t)b * nnzA);
    memcpy(sol_x, x + (size_t)b * %(n)d, sizeof(double) * %(n)d);
    memcpy(sol_y, y + (size_t)b * %(m)d, sizeof(double) * %(m)d);
    for (i = 0; i < %(n)d; i++) CPG_OSQP_Grad.dx[i] = dx[(size_t)b * %(n)d + i];
    cpg_P_to_K(&Pm, CPG_OSQP_Grad.K, CPG_OSQP_Grad.K_true);
    cpg_A_to_K(&Am, CPG_OSQP_Grad.K, CPG_OSQP_Grad.K_true);
    if (CPG_OSQP_Grad.init) {
      cpg_ldl_symbolic();
      cpg_ldl_numeric();
      for (j = 0; j < %(N)d - 1; j++)
        for (k = CPG_OSQP_Grad.L->p[j]; k < CPG_OSQP_Grad.L->p[j + 1]; k++) {
          i = CPG_OSQP_Grad.L->i[k];
          CPG_OSQP_Grad.Lmask[(2 * %(N)d - 3 - j) * j / 2 + i - 1] = 1;
        }
      CPG_OSQP_Grad.init = 0;
    } else {
      cpg_ldl_numeric();
      for (i = 0; i < %(m)d; i++) CPG_OSQP_Grad.a[i] = 1;
    }
    cpg_osqp_gradient();
    memcpy(dq + (size_t)b * %(n)d, CPG_OSQP_Grad.dq, sizeof(double) * %(n)d);
    memcpy(dl + (size_t)b * %(m)d, CPG_OSQP_Grad.dl, sizeof(double) * %(m)d);
    memcpy(du + (size_t)b * %(m)d, CPG_OSQP_Grad.du, sizeof(double) * %(m)d);
    memcpy(dP + (size_t)b * nnzP, CPG_OSQP_Grad.dP->x, sizeof(double) * nnzP);
    memcpy(dA + (size_t)b * nnzA, CPG_OSQP_Grad.dA->x, sizeof(double) * nnzA);
  }
}
'''


def build(name: str, P, A, out_dir=None) -> str:
    """P upper-triangular CSC, A CSC: the family's UNSCALED canonical matrices at its default parameters."""
    U = _ref_utils()
    P = sp.csc_matrix(P); A = sp.csc_matrix(A)
    Pfull = sp.csc_matrix(sp.triu(P) + sp.triu(P, 1).T)
    n, m = P.shape[0], A.shape[0]
    N = n + m
    out_dir = out_dir or os.path.join(HERE, '_ref')
    os.makedirs(out_dir, exist_ok=True)
    with tempfile.TemporaryDirectory() as td:
        ctx = {'n': n, 'N': N, 'workspace': 'CPG_OSQP_Grad', 'gradient_two_stage': False,
               'sol_x_var': 'sol_x', 'sol_y_var': 'sol_y'}
        for tpl, c in (('cpg_osqp_grad_compute.c.jinja2', ctx), ('cpg_osqp_grad_compute.h.jinja2', {}),
                       ('cpg_osqp_grad_workspace.h.jinja2', {'workspace': 'CPG_OSQP_Grad'})):
            U.render_template_to_file(tpl, td, c)
        with open(os.path.join(td, 'cpg_workspace.h'), 'w') as f:
            f.write('#ifndef CPG_TYPES_H\n#define CPG_TYPES_H\ntypedef double cpg_float;\ntypedef int cpg_int;\n'
                    'extern cpg_float sol_x[];\nextern cpg_float sol_y[];\n#endif\n')
        with open(os.path.join(td, 'qdldl_types.h'), 'w') as f:
            f.write(open(os.path.join(HERE, 'config', 'qdldl_types.h')).read())
        # static workspace: entry list of cvxpygen/writer.py:375-403
        K = sp.bmat([[Pfull + 1e-6 * sp.eye(n), A.T], [None, -1e-6 * sp.eye(m)]], format='csc')
        K = sp.triu(K, format='csc')          # the reference passes the upper triangle to QDLDL (P is stored upper)
        K_true = sp.bmat([[Pfull, A.T], [A, None]], format='csr')
        entries = [('a', 'int', np.ones(m, dtype=int)), ('etree', 'int', np.zeros(N, dtype=int)),
                   ('Lnz', 'int', np.zeros(N, dtype=int)), ('iwork', 'int', np.zeros(3 * N, dtype=int)),
                   ('bwork', 'int', np.zeros(N, dtype=int)), ('fwork', 'float', np.zeros(N)), ('L', 'csc_L', N),
                   ('Lmask', 'int', np.zeros((N - 1) * N // 2, dtype=int)), ('D', 'float', np.ones(N)),
                   ('Dinv', 'float', np.ones(N)), ('K', 'csc', K), ('K_true', 'csc', K_true),
                   ('rhs', 'float', np.zeros(N)), ('delta', 'float', np.zeros(N)), ('c', 'float', np.zeros(N)),
                   ('w', 'float', np.zeros(N)), ('wi', 'int', np.arange(N)), ('l', 'float', np.zeros(N)),
                   ('li', 'int', np.arange(N)), ('lx', 'float', np.zeros(N)), ('dx', 'float', np.zeros(n)),
                   ('r', 'float', np.zeros(N)), ('dq', 'float', np.zeros(n)), ('dl', 'float', np.zeros(m)),
                   ('du', 'float', np.zeros(m)), ('dP', 'csc', 0 * sp.csc_matrix(P)), ('dA', 'csc', 0 * A)]
        with open(os.path.join(td, 'cpg_osqp_grad_workspace.c'), 'w') as f:
            f.write('#include "cpg_osqp_grad_workspace.h"\n\n')
            for nm, typ, val in entries:
                if typ == 'csc':
                    U.write_mat_def(f, val, f'cpg_osqp_grad_{nm}', qualifier='grad')
                elif typ == 'csc_L':
                    U.write_L_def(f, val, f'cpg_osqp_grad_{nm}', qualifier='grad')
                else:
                    U.write_vec_def(f, val, f'cpg_osqp_grad_{nm}', 'cpg_' + typ, qualifier='grad')
            fields = ['init'] + [e[0] for e in entries]
            casts = [''] + [f'{U.type_to_cast(e[1], qualifier="grad")}&' for e in entries]
            values = ['1'] + [f'cpg_osqp_grad_{v}' for v in fields[1:]]
            U.write_struct_def(f, fields, casts, values, 'CPG_OSQP_Grad', 'CPG_OSQP_Grad_t')
        with open(os.path.join(td, 'driver.c'), 'w') as f:
            f.write(DRIVER % dict(n=n, m=m, N=N))
        so = os.path.join(out_dir, f'libgrad_ref_{name}.so')
        cmd = ['gcc', '-O2', '-fPIC', '-shared', '-w', '-I', td, '-I', os.path.join(QDLDL, 'include'),
               os.path.join(td, 'cpg_osqp_grad_compute.c'), os.path.join(td, 'cpg_osqp_grad_workspace.c'),
               os.path.join(td, 'driver.c'), os.path.join(QDLDL, 'src', 'qdldl.c'), '-lm', '-o', so]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode:
            raise RuntimeError(res.stderr[-3000:])
    return so


def grad_ref_batch(so, n, m, x, y, dx):
    import ctypes as C
    lib = C.CDLL(so)
    x = np.ascontiguousarray(x, dtype=np.float64); y = np.ascontiguousarray(y, dtype=np.float64)
    dx = np.ascontiguousarray(dx, dtype=np.float64)
    B = x.shape[0]
    dq = np.zeros((B, n)); dl = np.zeros((B, m)); du = np.zeros((B, m))
    p = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    lib.grad_ref_batch(C.c_int(B), p(x), p(y), p(dx), p(dq), p(dl), p(du))
    return dq, dl, du


def grad_ref_batch_mat(so, n, m, Px, Ax, x, y, dx):
    import ctypes as C
    lib = C.CDLL(so)
    c = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    Px, Ax, x, y, dx = c(Px), c(Ax), c(x), c(y), c(dx)
    B = x.shape[0]
    dq = np.zeros((B, n)); dl = np.zeros((B, m)); du = np.zeros((B, m))
    dP = np.zeros((B, Px.shape[1])); dA = np.zeros((B, Ax.shape[1]))
    p = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    lib.grad_ref_batch_mat(C.c_int(B), p(Px), p(Ax), p(x), p(y), p(dx), p(dq), p(dl), p(du), p(dP), p(dA))
    return dq, dl, du, dP, dA


def structural(M):
    """same pattern, every stored entry = 1: keeps structural zeros through scipy's sparse arithmetic."""
    M = sp.csc_matrix(M)
    return sp.csc_matrix((np.ones(len(M.indices)), M.indices.copy(), M.indptr.copy()), shape=M.shape)


if __name__ == '__main__':
    sys.path.insert(0, ROOT)
    from cvxpygen_b200 import standard
    for name in (sys.argv[1:] or ['mpc_6_3_10', 'nonneg_LS_3_2', 'random_qp_20_5_15', 'mpc_12_4_10']):
        fam = standard.STANDARD[name][0]()
        if name in standard.MATPAR_NAMES:     # every structural entry must exist in K / K_true: values are set per instance
            print('built', build(name, structural(fam.canon_matrix('P')), structural(fam.canon_matrix('A'))))
        else:
            print('built', build(name, fam.canon_matrix('P'), fam.canon_matrix('A')))
