"""The warp-per-instance CUDA kernels, run on the CPU.

tests/emu/simt/cuda_runtime.h is a small SIMT emulator (every CUDA thread a fiber, switches at the warp / block
synchronisation points); tests/emu/admm_emu.cpp builds the PRODUCT kernel sources of one generated family against it
-- admm_matpar_kernel (per-instance matrices), admm_tail_kernel (per-instance factor), qp_grad_kernel (backward pass).
The lane-level logic of these kernels (shuffle sweeps, table-driven factorisation, gather tables, in-warp equilibration)
is thereby checked against the oracle in the CPU-only suite, before any GPU time is spent; the GPU tests
(tests/test_matpar.py, tests/test_gpu_parity.py, tests/test_grad.py) remain the parity tests proper."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

from cvxpygen_b200 import codegen, families
from cvxpygen_b200.offline.qp_setup import setup_qp_family
from helpers import canon_batches, canon_matrix_batches, matrix_oracle_solve, oracle_solve, rel_err
from oracle.grad_numpy import qp_backward, qp_backward_mat, param_gradient, param_gradient_mat

HERE = os.path.dirname(os.path.abspath(__file__))


def build_emu(fam, batch, out_dir, flags=(), dmma=None, force_big=False):
    st = setup_qp_family(fam, batch)
    codegen.write_code(st, out_dir, dmma=dmma, force_big=force_big)
    inc, sol, src = (os.path.join(out_dir, 'c', d) for d in ('include', 'solver_code', 'src'))
    so = os.path.join(out_dir, 'libadmm_emu.so')
    cmd = ['g++', '-O2', '-std=c++17', '-fPIC', '-shared', '-DCPG_SIMT_HOST_EMU', '-w', '-ffp-contract=off', *flags,
           '-I', os.path.join(HERE, 'emu', 'simt'), '-I', inc, '-I', sol, os.path.join(HERE, 'emu', 'admm_emu.cpp'),
           '-x', 'c', os.path.join(src, 'cpg_blob.c'), '-o', so]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-3000:]
    lib = C.CDLL(so)
    dims = (C.c_int * 6)()
    lib.emu_dims(dims)
    return st, lib, list(dims)


def _ptr(a, t=C.c_double):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def run_solve(lib, fn, dims, rows, grid=2, adaptive_rho_interval=0, eps=1e-3):
    n, m, npb, n_prim, n_dual, _ = dims
    rows = np.ascontiguousarray(rows, dtype=np.float64)
    B = rows.shape[0]
    out = dict(prim=np.zeros((B, n_prim)), dual=np.zeros((B, n_dual)), x=np.zeros((B, n)), y=np.zeros((B, m)), obj=np.zeros(B),
               iter=np.zeros(B, np.int32), status=np.zeros(B, np.int32), pri=np.zeros(B), dua=np.zeros(B))
    rc = getattr(lib, fn)(C.c_int(B), _ptr(rows), _ptr(out['prim']), _ptr(out['dual']), _ptr(out['x']), _ptr(out['y']),
                          _ptr(out['obj']), _ptr(out['iter'], C.c_int), _ptr(out['status'], C.c_int), _ptr(out['pri']),
                          _ptr(out['dua']), C.c_int(grid), C.c_int(adaptive_rho_interval), C.c_double(eps))
    assert rc >= 0
    out['rc'] = rc
    return out


def _rows(fam, st, params, B):
    th = np.tile(fam.theta_default(), (B, 1))
    for pn, v in params.items():
        p = fam.param(pn)
        th[:, p.col:p.col + p.size] = v
    return th[:, st.batch_cols]


def test_matrix_parameter_kernel_on_the_emulator(tmp_path):
    """admm_matpar_kernel end to end (canonicalise, equilibrate, assemble, factor, ADMM with a rho update) for a small LTV MPC
    family, 7 instances over 2 blocks x 8 warps, against the reference / the numpy restatement of osqp_update_P_A + solve."""
    fam = families.mpc_ltv(4, 2, 5)
    batch = ['A', 'B', 'qdiag', 'rdiag', 'x_init']
    st, lib, dims = build_emu(fam, batch, str(tmp_path))
    assert dims[5] == 1
    B = 7
    params = families.mpc_ltv_batch(fam, B, seed=3)
    out = run_solve(lib, 'emu_matpar_solve', dims, _rows(fam, st, params, B))
    Px, Ax, (q, l, u) = canon_matrix_batches(fam, params, B)
    ora = matrix_oracle_solve(fam, Px, Ax, q, l, u)
    assert np.array_equal(out['iter'], ora['iter']) and np.array_equal(out['status'], ora['status'])
    assert rel_err(out['x'], ora['x']).max() < 1e-9 and rel_err(out['y'], ora['y']).max() < 1e-9
    assert np.allclose(out['obj'], ora['obj'], rtol=1e-9) and np.allclose(out['pri'], ora['pri_res'], rtol=1e-6, atol=1e-12)
    prim_idx = np.concatenate([v.indices for v in fam.variables])
    assert np.array_equal(out['prim'], out['x'][:, prim_idx])
    # tighter tolerance + rho adaptation every 25 iterations: in-place re-factorisations of the per-instance K
    kw = dict(adaptive_rho_interval=25, eps_abs=1e-6, eps_rel=1e-6)
    out2 = run_solve(lib, 'emu_matpar_solve', dims, _rows(fam, st, params, B), adaptive_rho_interval=25, eps=1e-6)
    ora2 = matrix_oracle_solve(fam, Px, Ax, q, l, u, **kw)
    assert ora2['rho_updates'].sum() > 0
    assert np.array_equal(out2['iter'], ora2['iter']) and rel_err(out2['x'], ora2['x']).max() < 1e-8


@pytest.mark.parametrize('dmma', [False, True])
def test_main_kernel_on_the_emulator(tmp_path, dmma):
    """The headline path: admm_multi_kernel (two instances per warp, generated straight-line KKT solve, lockstep CTA,
    persistent slots refilled from the work counter) followed by admm_tail_kernel on the instances it hands off; and its
    tensor-core variant admm_dmma_kernel (groups of four warps, mma.m8n8k4.f64 emulated lane by lane, named barriers)."""
    fam = families.mpc(4, 2, 6)
    st, lib, dims = build_emu(fam, ['x_init'], str(tmp_path), dmma=dmma)
    B = 29                                         # odd: the last warp slot runs half empty
    xi = np.random.default_rng(11).uniform(-1.5, 1.5, (B, 4))
    q, l, u = canon_batches(fam, {'x_init': xi}, B)
    out = run_solve(lib, 'emu_main_solve', dims, xi, grid=2)
    ora = oracle_solve(fam, q, l, u)
    assert (out['status'] != -100).all()
    assert np.array_equal(out['iter'], ora['iter']) and np.array_equal(out['status'], ora['status'])
    assert rel_err(out['x'], ora['x']).max() < 1e-9 and rel_err(out['y'], ora['y']).max() < 1e-9
    assert np.allclose(out['obj'], ora['obj'], rtol=1e-9)
    prim_idx = np.concatenate([v.indices for v in fam.variables]); dual_idx = np.concatenate([d.indices for d in fam.duals])
    assert np.array_equal(out['prim'], out['x'][:, prim_idx]) and np.array_equal(out['dual'], out['y'][:, dual_idx])
    one = run_solve(lib, 'emu_main_solve', dims, xi[:1], grid=1)           # a single instance: one half-filled warp slot
    assert one['iter'][0] == ora['iter'][0] and np.array_equal(one['x'][0], out['x'][0])
    # rho adaptation every 25 iterations at 1e-6: hand-offs to the tail kernel
    kw = dict(adaptive_rho_interval=25, eps_abs=1e-6, eps_rel=1e-6)
    out2 = run_solve(lib, 'emu_main_solve', dims, xi, grid=2, adaptive_rho_interval=25, eps=1e-6)
    ora2 = oracle_solve(fam, q, l, u, **kw)
    assert ora2['rho_updates'].sum() > 0 and (out2['status'] != -100).all()
    assert np.array_equal(out2['iter'], ora2['iter']) and rel_err(out2['x'], ora2['x']).max() < 1e-8


def test_main_kernel_warm_start_on_the_emulator(tmp_path):
    """Warm start (osqp_warm_start, osqp.c:929-953) through the main kernel: its per-instance state is the pre-projection vector
    t, which cannot represent a start point (z0 = A x0 need not lie in [l, u]); the first iteration reads (z0, y0 / rho) from the
    instance's scratch rows instead.  Same iteration counts as the reference warm-started from the same point."""
    fam = families.mpc(4, 2, 6)
    st, lib, dims = build_emu(fam, ['x_init'], str(tmp_path))
    n, m = dims[0], dims[1]
    B = 13
    rng = np.random.default_rng(5)
    xi = rng.uniform(-1.5, 1.5, (B, 4))
    q, l, u = canon_batches(fam, {'x_init': xi}, B)
    cold = oracle_solve(fam, q, l, u)
    # start from the solution of a neighbouring instance, perturbed: z0 = A x0 violates the bounds of most rows
    x0 = np.ascontiguousarray(np.roll(cold['x'], 1, axis=0) + 0.05 * rng.standard_normal((B, n)))
    y0 = np.ascontiguousarray(np.roll(cold['y'], 1, axis=0) + 0.05 * rng.standard_normal((B, m)))
    out = dict(prim=np.zeros((B, dims[3])), dual=np.zeros((B, dims[4])), x=np.zeros((B, n)), y=np.zeros((B, m)), obj=np.zeros(B),
               iter=np.zeros(B, np.int32), status=np.zeros(B, np.int32), pri=np.zeros(B), dua=np.zeros(B))
    rows = np.ascontiguousarray(xi)
    rc = lib.emu_main_solve_warm(C.c_int(B), _ptr(rows), _ptr(x0), _ptr(y0), _ptr(out['prim']), _ptr(out['dual']), _ptr(out['x']),
                                 _ptr(out['y']), _ptr(out['obj']), _ptr(out['iter'], C.c_int), _ptr(out['status'], C.c_int),
                                 _ptr(out['pri']), _ptr(out['dua']), C.c_int(2), C.c_int(0), C.c_double(1e-3))
    assert rc >= 0
    ora = oracle_solve(fam, q, l, u, x0=x0, y0=y0)
    assert np.array_equal(out['iter'], ora['iter']) and np.array_equal(out['status'], ora['status'])
    assert rel_err(out['x'], ora['x']).max() < 1e-9 and rel_err(out['y'], ora['y']).max() < 1e-9
    assert not np.array_equal(ora['iter'], cold['iter'])          # the start point mattered


@pytest.mark.parametrize('name,B,dmma', [('nonneg_LS_3_2', 40, False), ('box_qp_6_8', 96, False), ('random_qp_20_5_15', 24, False),
                                         ('mpc_12_4_10', 12, False), ('portfolio_qp_50_10', 6, False),
                                         ('box_qp_6_8', 96, True)])      # (the tensor-core variant on mpc: test_main_kernel_on_the_emulator[True] and the GPU suite)
def test_standard_families_on_the_emulator(name, B, dmma, tmp_path):
    """Main + tail kernels of the standard families -- incl. the headline MPC-12/4/10 family with its 374-step generated solve and
    the 742-row portfolio QP --: unstructured sparsity with q, l, u all batched; box_qp's corner
    cases -- bounds that change a constraint's type (hand-off at iteration 0), primal and dual infeasibility certificates,
    no-solution statuses with NaN solutions and +-1e30 objectives."""
    from cvxpygen_b200 import standard
    from helpers import family_and_batch, assert_batch_parity, rounding_stable
    from types import SimpleNamespace
    fam, params, (q, l, u) = family_and_batch(name, B, seed=12)
    if name == 'box_qp_6_8':            # regular / equality-collapsed / loose rows, primal infeasible, dual infeasible
        from test_gpu_parity import corner_case_batch
        params, kind = corner_case_batch(fam, B)
        q, l, u = canon_batches(fam, params, B)
    batch = standard.STANDARD[name][1]
    st, lib, dims = build_emu(fam, batch, str(tmp_path), dmma=dmma)      # dmma: the tensor-core variant of the main kernel
    out = run_solve(lib, 'emu_main_solve', dims, _rows(fam, st, params, B), grid=2)
    ora = oracle_solve(fam, q, l, u)
    info = SimpleNamespace(status=out['status'], iter=out['iter'], obj_val=out['obj'], pri_res=out['pri'], dua_res=out['dua'])
    stable = rounding_stable(fam, q, l, u, ora) if name == 'box_qp_6_8' else None
    assert_batch_parity(out['x'], out['y'], info, ora, 1e-8, stable=stable, obj_sign=-1.0 if fam.is_maximization else 1.0)
    if name == 'box_qp_6_8':
        assert out['rc'] > 0                                             # some instances went through the tail kernel
        assert set(np.unique(ora['status'])) >= {1, -3, -4} and np.array_equal(out['status'], ora['status'])
        assert (out['obj'][kind == 3] == 1e30).all() and (out['obj'][kind == 4] == -1e30).all()


def test_big_family_path_on_the_emulator(tmp_path):
    """Families whose tile schedule exceeds shared memory (forced here on a small one): no main kernel, every instance queued at
    iteration 0 for the per-instance-factor kernel, which reads the constants from global memory instead of staging them."""
    fam = families.mpc(4, 2, 6)
    st, lib, dims = build_emu(fam, ['x_init'], str(tmp_path), force_big=True)
    hdr = open(os.path.join(str(tmp_path), 'c', 'include', 'cpg_family.h')).read()
    assert '#define CPG_FAM_BIG 1' in hdr and '#define CPG_FAM_TAIL_STAGE 0' in hdr
    B = 11
    xi = np.random.default_rng(3).uniform(-1.5, 1.5, (B, 4))
    q, l, u = canon_batches(fam, {'x_init': xi}, B)
    out = run_solve(lib, 'emu_main_solve', dims, xi, grid=2)
    ora = oracle_solve(fam, q, l, u)
    assert out['rc'] == B and (out['status'] != -100).all()            # all of them went through the queue
    assert np.array_equal(out['iter'], ora['iter']) and np.array_equal(out['status'], ora['status'])
    assert rel_err(out['x'], ora['x']).max() < 1e-9 and rel_err(out['y'], ora['y']).max() < 1e-9


def test_tail_kernel_on_the_emulator(tmp_path):
    """admm_tail_kernel: every instance queued at iteration 0 (the route of a constraint-type change), so the kernel factors
    K numerically on the symbolic pattern (dense-group packed triangles included) and runs the whole ADMM loop on it."""
    fam = families.mpc(4, 2, 6)
    st, lib, dims = build_emu(fam, ['x_init'], str(tmp_path))
    B = 6
    xi = np.random.default_rng(4).uniform(-1, 1, (B, 4))
    q, l, u = canon_batches(fam, {'x_init': xi}, B)
    for kw, ari, eps in ((dict(), 0, 1e-3), (dict(adaptive_rho_interval=25, eps_abs=1e-6, eps_rel=1e-6), 25, 1e-6)):
        out = run_solve(lib, 'emu_tail_solve', dims, xi, adaptive_rho_interval=ari, eps=eps)
        ora = oracle_solve(fam, q, l, u, **kw)
        assert np.array_equal(out['iter'], ora['iter']) and np.array_equal(out['status'], ora['status'])
        assert rel_err(out['x'], ora['x']).max() < 1e-9 and rel_err(out['y'], ora['y']).max() < 1e-9


def _run_grad(lib, dims, rows, sol_x, sol_y, dprim, nnzP=0, nnzA=0, grid=2):
    n, m, npb, n_prim, n_dual, matpar = dims
    B = sol_y.shape[0]
    c = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)
    rows, sol_x, sol_y, dprim = c(rows), c(sol_x), c(sol_y), c(dprim)
    out = dict(dparams=np.zeros((B, max(npb, 1))), dq=np.zeros((B, n)), dl=np.zeros((B, m)), du=np.zeros((B, m)),
               dP=np.zeros((B, max(nnzP, 1))), dA=np.zeros((B, max(nnzA, 1))))
    rc = lib.emu_gradient(C.c_int(B), _ptr(rows), _ptr(sol_x), _ptr(sol_y), _ptr(dprim), _ptr(out['dparams']), _ptr(out['dq']),
                          _ptr(out['dl']), _ptr(out['du']), _ptr(out['dP']) if matpar else None, _ptr(out['dA']) if matpar else None,
                          C.c_int(grid))
    assert rc == 0
    return out


def test_backward_kernels_on_the_emulator(tmp_path):
    """qp_grad_kernel (shared matrices) and qp_grad_kernel<Fam, true> (per-instance matrices, dP / dA folded into dtheta)
    against the numpy restatement of the reference's cpg_osqp_gradient."""
    # shared matrices
    fam = families.mpc(4, 2, 6)
    st, lib, dims = build_emu(fam, ['x_init'], str(tmp_path / 'a'))
    B = 5
    xi = np.random.default_rng(6).uniform(-1, 1, (B, 4))
    q, l, u = canon_batches(fam, {'x_init': xi}, B)
    sol = oracle_solve(fam, q, l, u, eps_abs=1e-8, eps_rel=1e-8)
    prim_idx = np.concatenate([v.indices for v in fam.variables])
    dprim = np.random.default_rng(7).standard_normal((B, len(prim_idx)))
    dx = np.zeros((B, fam.n_var)); dx[:, prim_idx] = dprim
    got = _run_grad(lib, dims, None, None, sol['y'], dprim)
    dq, dl, du, _ = qp_backward(fam.canon_matrix('P'), fam.canon_matrix('A'), sol['x'], sol['y'], dx)
    assert np.abs(got['dq'] - dq).max() < 1e-7 * np.abs(dq).max() and np.abs(got['dl'] + got['du'] - dl - du).max() < 1e-7 * np.abs(dl + du).max()
    want = param_gradient(fam, dq, dl, du, ['x_init'])
    assert np.abs(got['dparams'][:, :4] - want).max() < 1e-7 * np.abs(want).max()
    # per-instance matrices
    fam = families.mpc_ltv(4, 2, 5)
    batch = ['A', 'B', 'qdiag', 'rdiag', 'x_init']
    st, lib, dims = build_emu(fam, batch, str(tmp_path / 'b'))
    params = families.mpc_ltv_batch(fam, B, seed=8)
    Px, Ax, (q, l, u) = canon_matrix_batches(fam, params, B)
    sol = matrix_oracle_solve(fam, Px, Ax, q, l, u, eps_abs=1e-8, eps_rel=1e-8)
    prim_idx = np.concatenate([v.indices for v in fam.variables])
    dprim = np.random.default_rng(9).standard_normal((B, len(prim_idx)))
    dx = np.zeros((B, fam.n_var)); dx[:, prim_idx] = dprim
    got = _run_grad(lib, dims, _rows(fam, st, params, B), sol['x'], sol['y'], dprim, nnzP=Px.shape[1], nnzA=Ax.shape[1])
    dq, dl, du, dP, dA = qp_backward_mat(fam.patterns['P'], fam.patterns['A'], Px, Ax, sol['x'], sol['y'], dx)
    for k, ref in (('dq', dq), ('dP', dP), ('dA', dA)):
        assert np.abs(got[k] - ref).max() < 1e-7 * max(np.abs(ref).max(), 1e-30), k
    want = param_gradient_mat(fam, dq, dl, du, dP, dA, batch)
    assert np.abs(got['dparams'] - want).max() < 1e-7 * np.abs(want).max()


def test_results_do_not_depend_on_the_thread_schedule(tmp_path):
    """Race check: the emulator resumes runnable threads in ascending, descending or shuffled order; a lane that consumed
    another lane's shared-memory value without a barrier in between would see a different value under a different order.
    The main kernel (no atomics on its path) must return bit-identical results under every schedule; the kernels that
    factor K numerically add their update terms with shared-memory atomics, whose ORDER is the schedule's -- there the
    results may differ in the last bits only (same iteration counts, 1e-12)."""
    fam = families.mpc(4, 2, 6)
    st, lib, dims = build_emu(fam, ['x_init'], str(tmp_path / 'mpc'))
    xi = np.random.default_rng(2).uniform(-1.5, 1.5, (21, 4))
    base = None
    for mode in (0, 1, 2):
        lib.emu_set_schedule(mode)
        out = run_solve(lib, 'emu_main_solve', dims, xi)
        assert out['rc'] == 0                                  # nothing handed off: the whole solve ran in admm_multi_kernel
        cur = (out['x'].tobytes(), out['y'].tobytes(), out['iter'].tobytes(), out['status'].tobytes(), out['obj'].tobytes())
        base = base or cur
        assert cur == base, f'schedule {mode} changes the result'
    lib.emu_set_schedule(0)
    fam = families.mpc_ltv(4, 2, 5)
    batch = ['A', 'B', 'qdiag', 'rdiag', 'x_init']
    st, lib, dims = build_emu(fam, batch, str(tmp_path / 'ltv'))
    B = 5
    params = families.mpc_ltv_batch(fam, B, seed=21)
    rows = _rows(fam, st, params, B)
    prim_idx = np.concatenate([v.indices for v in fam.variables])
    dprim = np.random.default_rng(1).standard_normal((B, len(prim_idx)))
    base = None
    for mode in (0, 1, 2):
        lib.emu_set_schedule(mode)
        out = run_solve(lib, 'emu_matpar_solve', dims, rows, adaptive_rho_interval=25, eps=1e-5)
        g = _run_grad(lib, dims, rows, out['x'], out['y'], dprim, nnzP=st.nnzP, nnzA=st.nnzA)
        if base is None:
            base = (out, g)
            continue
        assert np.array_equal(out['iter'], base[0]['iter']) and np.array_equal(out['status'], base[0]['status'])
        assert rel_err(out['x'], base[0]['x']).max() < 1e-12 and rel_err(out['y'], base[0]['y']).max() < 1e-12
        assert np.abs(g['dparams'] - base[1]['dparams']).max() < 1e-10 * np.abs(base[1]['dparams']).max()
    lib.emu_set_schedule(0)


@pytest.mark.parametrize('form', [2, 1])
def test_atomics_free_factorisation_is_deterministic(form, tmp_path):
    """tail_factor without atomics -- form 2, the default: coloured rounds (32 ops with pairwise distinct targets per round, plain
    read-modify-writes); form 1: owner-writes (one lane sums a target's ops) -- gives the reference's answers and, with no atomic
    left on the path, bit-identical results under every thread schedule, for the solve and for the backward pass."""
    fam = families.mpc_ltv(4, 2, 5)
    batch = ['A', 'B', 'qdiag', 'rdiag', 'x_init']
    st, lib, dims = build_emu(fam, batch, str(tmp_path), flags=(f'-DCPG_TAIL_FACTOR_FORM={form}',))
    B = 6
    params = families.mpc_ltv_batch(fam, B, seed=33)
    rows = _rows(fam, st, params, B)
    Px, Ax, (q, l, u) = canon_matrix_batches(fam, params, B)
    kw = dict(adaptive_rho_interval=25, eps_abs=1e-6, eps_rel=1e-6)
    ora = matrix_oracle_solve(fam, Px, Ax, q, l, u, **kw)
    prim_idx = np.concatenate([v.indices for v in fam.variables])
    dprim = np.random.default_rng(3).standard_normal((B, len(prim_idx)))
    base = None
    for mode in (0, 1, 2):
        lib.emu_set_schedule(mode)
        out = run_solve(lib, 'emu_matpar_solve', dims, rows, adaptive_rho_interval=25, eps=1e-6)
        g = _run_grad(lib, dims, rows, out['x'], out['y'], dprim, nnzP=st.nnzP, nnzA=st.nnzA)
        cur = (out['x'].tobytes(), out['y'].tobytes(), out['iter'].tobytes(), g['dparams'].tobytes(), g['dA'].tobytes())
        if base is None:
            base = cur
            assert np.array_equal(out['iter'], ora['iter']) and rel_err(out['x'], ora['x']).max() < 1e-8
        assert cur == base, f'schedule {mode} changes the result'
    lib.emu_set_schedule(0)
    # the tables themselves: owner-writes == push form up to rounding
    from cvxpygen_b200.offline import kkt, refactor
    rv = kkt.rho_vector(st.ctype, 0.37)
    a, b, c = (f(st.refactor, rv) for f in (refactor.emulate_factor, refactor.emulate_factor_gather, refactor.emulate_factor_coloured))
    assert np.abs(a - b).max() < 1e-13 * np.abs(a).max() and np.abs(a - c).max() < 1e-13 * np.abs(a).max()
    nr = st.refactor.c_round_ptr[-1]
    assert nr * 32 >= len(st.refactor.ops) and nr <= sum(-(-(b_ - a_) // 32) for a_, b_ in zip(st.refactor.op_ptr[:-1], st.refactor.op_ptr[1:])) * 1.25


# ---------------------------------------------------------------------------------------------------------------------
# interior-point kernel: the DEVICE code path of csrc/ipm_kernel.cuh (not the CPG_IPM_HOST_EMU branches) on the emulator
def _build_ipm_simt(st, d):
    from cvxpygen_b200 import codegen_ipm
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, 'cpg_ipm_family.h'), 'w') as f:
        f.write(codegen_ipm.family_header(st))
    so = os.path.join(d, 'libipm_simt.so')
    res = subprocess.run(['g++', '-O2', '-std=c++17', '-shared', '-fPIC', '-w', '-DCPG_SIMT_HOST_EMU', '-ffp-contract=off',
                          '-I', os.path.join(HERE, 'emu', 'simt'), '-I', d, '-I', os.path.join(os.path.dirname(HERE), 'cvxpygen_b200', 'csrc'),
                          os.path.join(HERE, 'emu', 'ipm_simt.cpp'), '-o', so], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-3000:]
    return C.CDLL(so)


def _ipm_simt_solve(lib, st, params, maxit=100, grid=1):
    D = st.defines
    params = np.ascontiguousarray(params, dtype=np.float64)
    B = params.shape[0]
    out = dict(prim=np.zeros((B, D['NPRIM'])), dual=np.zeros((B, D['NDUAL'])), x=np.zeros((B, D['N'])), y=np.zeros((B, max(D['P'], 1))),
               z=np.zeros((B, D['M'])), s=np.zeros((B, D['M'])), obj=np.zeros(B), iter=np.zeros(B, np.int32),
               status=np.zeros(B, np.int32), pres=np.zeros(B), dres=np.zeros(B))
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    lib.ipm_simt_solve(st.smem_blob, st.gmem_blob, B, P(params), *[P(out[k]) for k in
                       ('prim', 'dual', 'x', 'y', 'z', 's', 'obj', 'iter', 'status', 'pres', 'dres')], maxit, grid)
    return out


@pytest.mark.parametrize('name,B', [('adp_socp_6_3', 2), ('network_lp_50_10', 2), ('portfolio_100_10', 1)])
def test_interior_point_device_code_on_the_emulator(name, B, tmp_path):
    """ipm_kernel as the GPU runs it -- one 256-thread CTA per instance, ~118 barrier-separated phases per iteration, gather
    plans finished with butterfly shuffles, REDUX max reductions, instances pulled from the work counter -- against the
    golden vectors of the compiled ECOS: identical exit flags and iteration counts, x / s to 1e-7.  portfolio_100_10 is
    BASELINE config 3's family."""
    from cvxpygen_b200.offline import socp_setup as ss
    from helpers import GOLDEN
    fam, batch = {'adp_socp_6_3': (families.adp_socp(), ['f']), 'network_lp_50_10': (families.network_lp(50, 10), ['c', 'w', 'f_min', 'f_max']),
                  'portfolio_100_10': (families.portfolio_socp(), ['a', 'w_prev'])}[name]
    g = np.load(os.path.join(GOLDEN, f'socp_{name}.npz'))
    st = ss.setup_socp_family(fam, batch)
    lib = _build_ipm_simt(st, str(tmp_path))
    P = np.concatenate([g['param_' + k][:B] for k in batch], axis=1)
    out = _ipm_simt_solve(lib, st, P, grid=2)
    assert np.array_equal(out['status'], g['exitflag'][:B]) and np.array_equal(out['iter'], g['iter'][:B])
    assert np.abs(out['x'] - g['x'][:B]).max() < 1e-7 * np.abs(g['x'][:B]).max()
    assert np.abs(out['s'] - g['s'][:B]).max() < 1e-7 * max(1.0, np.abs(g['s'][:B]).max())
    sign = -1.0 if fam.is_maximization else 1.0
    assert np.allclose(out['obj'], sign * g['pcost'][:B], rtol=1e-8, atol=1e-9)


def test_interior_point_kernel_is_schedule_independent(tmp_path):
    """Race check for the kernel with the most barriers: bit-identical iterates and iteration counts whether the emulator resumes
    the 256 threads of the CTA in ascending, descending or shuffled order (a phase reading what another thread writes in the
    same phase would break this); DESIGN claims bitwise reproducibility for this kernel -- no atomics, fixed summation order."""
    from cvxpygen_b200.offline import socp_setup as ss
    from helpers import GOLDEN
    fam = families.adp_socp()
    g = np.load(os.path.join(GOLDEN, 'socp_adp_socp_6_3.npz'))
    st = ss.setup_socp_family(fam, ['f'])
    lib = _build_ipm_simt(st, str(tmp_path))
    base = None
    for mode in (0, 1, 2):
        lib.ipm_simt_set_schedule(mode)
        out = _ipm_simt_solve(lib, st, g['param_f'][:1])
        cur = tuple(out[k].tobytes() for k in ('x', 'z', 's', 'iter', 'status', 'obj'))
        base = base or cur
        assert cur == base, f'schedule {mode} changes the result'
    lib.ipm_simt_set_schedule(0)


def test_interior_point_matrix_parameters_on_the_emulator(tmp_path):
    """Device code path of ipm_kernel with per-instance G / A values (IPM_MATPAR): entries from the parameter row, three
    equilibration passes (atomic maxima on shared memory, cone sums on one thread), K rebuilt from the instance's entries every
    iteration, output scalings in the CTA's scratch -- against ECOS_updateData + ECOS_solve of the compiled reference."""
    from cvxpygen_b200.offline import socp_setup as ss
    from oracle import ref_ecos
    if not ref_ecos.available():
        pytest.skip('oracle/_ref/libecos_ref.so not built')
    sys.path.insert(0, HERE)
    from test_socp_ipm import _portfolio_mat_batch, _conic_reference_mat
    fam = families.portfolio_socp(12, 3, matrix_params=True)
    names = ['a', 'w_prev', 'F', 'd_sqrt']
    st = ss.setup_socp_family(fam, names)
    lib = _build_ipm_simt(st, str(tmp_path))
    B = 3
    params = _portfolio_mat_batch(fam, B, seed=2)
    out = _ipm_simt_solve(lib, st, np.concatenate([params[k] for k in names], axis=1), grid=2)
    ref = _conic_reference_mat(fam, params, B)
    assert np.array_equal(out['status'], ref['exitflag']) and np.array_equal(out['iter'], ref['iter'])
    assert np.abs(out['x'] - ref['x']).max() < 1e-7 * np.abs(ref['x']).max()
