#!/usr/bin/env python
"""Golden vectors of the backward pass, produced by the REFERENCE's own generated C (cpg_osqp_gradient, rendered from
cvxpygen/templates/cpg_osqp_grad_compute.c.jinja2 and compiled by oracle/build_grad_ref.py) on forward solutions of the
unmodified reference OSQP (oracle/_ref/libosqp_ref.so).  Run in the build container:

    make -C oracle ref && python oracle/build_grad_ref.py && python tests/golden/make_golden_grad.py

Stored per family: batched parameters, canonical forward solution (x, y) at cvxpygen's default settings, the upstream
gradient dprim on the user variables, and the reference's dq, dl, du.  (dl/du individually depend on the reference's
call history -- see csrc/grad_kernel.cuh -- so tests compare dq, dl + du and the parameter gradient.)"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from helpers import family_and_batch, oracle_solve          # noqa: E402
from oracle.build_grad_ref import grad_ref_batch, build      # noqa: E402
from cvxpygen_b200 import standard                           # noqa: E402

B = 48
for name in ('mpc_12_4_10', 'mpc_6_3_10', 'nonneg_LS_3_2'):
    fam, params, (q, l, u) = family_and_batch(name, B, seed=77)
    so = os.path.join(os.path.dirname(os.path.dirname(HERE)), 'oracle', '_ref', f'libgrad_ref_{name}.so')
    if not os.path.exists(so):
        build(name, fam.canon_matrix('P'), fam.canon_matrix('A'))
    sol = oracle_solve(fam, q, l, u)
    n, m = fam.n_var, fam.n_eq + fam.n_ineq
    prim_idx = np.concatenate([v.indices for v in fam.variables])
    dprim = np.random.default_rng(5).standard_normal((B, len(prim_idx)))
    dx = np.zeros((B, n)); dx[:, prim_idx] = dprim
    dq, dl, du = grad_ref_batch(so, n, m, sol['x'], sol['y'], dx)
    out = dict(sol_x=sol['x'], sol_y=sol['y'], dprim=dprim, dq=dq, dl=dl, du=du)
    for k, v in params.items():
        out['param_' + k] = v
    np.savez_compressed(os.path.join(HERE, f'grad_{name}.npz'), **out)
    print(name, 'active fraction', float((np.abs(sol['y']) > 1e-12).mean()), '|dq|max', float(np.abs(dq).max()))

# ---- families with per-instance MATRIX parameters (rows a16 + f2): the reference's P / A branch of cpg_gradient
# (cpg_P_to_K, cpg_A_to_K, cpg_ldl_numeric, cpg_osqp_gradient; cvxpygen/writer.py:240-263) incl. dP and dA
from helpers import ltv_batch, canon_matrix_batches, matrix_oracle_solve     # noqa: E402
from oracle.build_grad_ref import grad_ref_batch_mat, structural              # noqa: E402

for name in ('mpc_ltv_6_3_10',):
    fam = standard.STANDARD[name][0]()
    Bm = 24
    params = ltv_batch(fam, Bm, seed=78)
    Px, Ax, (q, l, u) = canon_matrix_batches(fam, params, Bm)
    so = os.path.join(os.path.dirname(os.path.dirname(HERE)), 'oracle', '_ref', f'libgrad_ref_{name}.so')
    if not os.path.exists(so):
        build(name, structural(fam.canon_matrix('P')), structural(fam.canon_matrix('A')))
    sol = matrix_oracle_solve(fam, Px, Ax, q, l, u)
    n, m = fam.n_var, fam.n_eq + fam.n_ineq
    prim_idx = np.concatenate([v.indices for v in fam.variables])
    dprim = np.random.default_rng(6).standard_normal((Bm, len(prim_idx)))
    dx = np.zeros((Bm, n)); dx[:, prim_idx] = dprim
    dq, dl, du, dP, dA = grad_ref_batch_mat(so, n, m, Px, Ax, sol['x'], sol['y'], dx)
    out = dict(sol_x=sol['x'], sol_y=sol['y'], dprim=dprim, dq=dq, dl=dl, du=du, dP=dP, dA=dA, Px=Px, Ax=Ax)
    for k, v in params.items():
        out['param_' + k] = v
    np.savez_compressed(os.path.join(HERE, f'grad_mat_{name}.npz'), **out)
    print(name, 'active fraction', float((np.abs(sol['y']) > 1e-12).mean()), '|dA|max', float(np.abs(dA).max()))
