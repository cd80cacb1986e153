"""BASELINE-size batches against the compiled unmodified reference, instance by instance (task 1a/1b).

Round 1 checked the full-size batches through size-independent properties only; `oracle/_ref` does ~80 k MPC inst/s and
~5.5 k SOCP inst/s on the GPU box's host threads, so comparing EVERY instance costs seconds:

  config 2  100 000 MPC QP instances        vs vendored OSQP 0.6.2 (osqp_update_bounds + osqp_solve)
  config 3   50 000 portfolio SOCP instances vs vendored ECOS 2.0.8 (ECOS_updateData + ECOS_solve), user-level primal AND dual
  config 4  100 000 backward passes          vs the reference's own generated gradient C (cpg_osqp_gradient)

Tolerance: 1e-5 relative per instance on primal / dual variables (BASELINE.json north_star).  Iteration counts must be identical
wherever the stopping test is not decided by the last bits (a different, equally exact KKT solve moves a residual that sits
on the threshold across it: such instances are counted, bounded, and must still agree to the solver's own tolerance)."""
import os

import numpy as np
import pytest

from cvxpygen_b200 import standard

import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

TOL = 1e-5
NT = os.cpu_count() or 1


def rel_rows(a, b):
    return np.linalg.norm(a - b, axis=1) / np.maximum(np.linalg.norm(b, axis=1), 1e-12)


@pytest.mark.gpu
def test_config2_full_batch_every_instance_vs_reference():
    wl = bench.WORKLOADS['mpc']
    B = 100000
    P = wl.host_params(B, 1)
    mod = standard.load(wl.family)
    res = mod.solve_batch(P, return_canonical=True)
    ora = wl.reference(P, NT)
    assert (ora['status'] == 1).all()
    assert np.array_equal(res.cpg_info.status, ora['status'])
    same = res.cpg_info.iter == ora['iter']
    # identical iteration counts except where the termination test is decided by rounding (count bounded: < 1 in 10 000)
    assert same.mean() >= 0.9999, f'{(~same).sum()} of {B} instances stop at a different check'
    ex, ey = rel_rows(res.sol_x, ora['x']), rel_rows(res.sol_y, ora['y'])
    assert ex[same].max() < TOL and ey[same].max() < TOL, (ex[same].max(), ey[same].max())
    if (~same).any():      # one check (25 iterations) apart: both are eps-solutions of the same instance
        assert np.abs(res.cpg_info.iter[~same] - ora['iter'][~same]).max() <= 25
        assert ex[~same].max() < 1e-2 and ey[~same].max() < 1e-2
    assert np.allclose(res.cpg_info.obj_val[same], ora['obj'][same], rtol=1e-6, atol=1e-9)
    # user-level retrieval of every instance = gather of its canonical solution (a12)
    fam = wl.canonical(P[:1])[0]
    prim_idx = np.concatenate([v.indices for v in fam.variables]); dual_idx = np.concatenate([d.indices for d in fam.duals])
    assert np.array_equal(res.prim, res.sol_x[:, prim_idx]) and np.array_equal(res.dual, res.sol_y[:, dual_idx])


@pytest.mark.gpu
def test_config3_full_batch_user_level_primal_and_dual_vs_reference():
    wl = bench.WORKLOADS['portfolio_socp']
    B = 50000
    P = wl.host_params(B, 1)
    mod = standard.load(wl.family)
    res = mod.solve_batch(P, return_canonical=True)
    fam, ora = wl.reference(P, NT)
    st, ef = res.cpg_info.status, ora['exitflag']
    # exit flags: 0 (optimal) everywhere except a handful of instances that ECOS itself -- or this kernel -- reports as 10
    # ("optimal inaccurate": the 1e-8 targets were not met before a numerical safeguard stopped the iteration, the 5e-5
    # `_inacc` tolerances are).  Which side of that decision an instance falls on is decided by rounding (measured: 4 of 50 000
    # differ, profiles/r2_socp_full_batch_diag.json); both sides must call every instance optimal to one of the two accuracies.
    assert np.isin(st, [0, 10]).all() and np.isin(ef, [0, 10]).all()
    assert (st == ef).mean() >= 0.9998, f'{(st != ef).sum()} exit flags differ'
    exact = (st == 0) & (ef == 0)
    prim_ref = np.concatenate([ora['x'][:, v.indices] for v in fam.variables], axis=1)
    dual_ref = np.concatenate([ora[d.vec][:, d.indices] for d in fam.duals], axis=1)
    ep, ed = rel_rows(res.prim, prim_ref), rel_rows(res.dual, dual_ref)
    # north_star: user-level primal AND dual within 1e-5 on EVERY instance both solvers call optimal ...
    assert ep[exact].max() < TOL, f'user-level primal: {ep[exact].max():.2e} at {ep.argmax()}'
    assert ed[exact].max() < TOL, f'user-level dual: {ed[exact].max():.2e} at {ed.argmax()}'
    # ... and within the reference's own "inaccurate" tolerance on the few where one of them stopped early
    assert ep.max() < 1e-4 and ed.max() < 1e-4
    # canonical x, y, s likewise; iteration counts identical except where the exit test at 1e-8 is decided by the last bits
    # (measured: 20 of 50 000 differ, by one iteration)
    for got, ref in ((res.sol_x, ora['x']), (res.sol_y, ora['y']), (res.sol_s, ora['s'])):
        e = rel_rows(got, ref)
        assert e[exact].max() < TOL and e.max() < 1e-4
    dit = np.abs(res.cpg_info.iter.astype(np.int64) - ora['iter'])
    assert dit.max() <= 1 and (dit == 0).mean() >= 0.999, (dit.max(), (dit == 0).mean())
    assert np.allclose(-res.cpg_info.obj_val, ora['pcost'], rtol=1e-7, atol=1e-9)      # maximisation: cpg reports -(pcost + d), d = 0


@pytest.mark.gpu
def test_config4_full_batch_backward_vs_reference_generated_c():
    wl = bench.WORKLOADS['mpc_grad']
    if not os.path.exists(os.path.join(bench.ROOT, 'oracle', '_ref', 'libgrad_ref_mpc_12_4_10.so')):
        pytest.skip('oracle/_ref/libgrad_ref_mpc_12_4_10.so not built')
    B = 100000
    P = wl.host_params(B, 1)
    mod = standard.load(wl.family)
    res = mod.solve_batch(P, return_canonical=True)
    dprim = np.random.default_rng(5).standard_normal((B, mod.dims.n_prim))
    got = mod.gradient_batch(res.sol_y, dprim)['x_init']
    fam = wl.canonical(P[:1])[0]
    ref, _ = wl.reference_backward(fam, res.sol_x, res.sol_y, dprim, NT)
    err = np.abs(got - ref).max(axis=1) / np.maximum(np.abs(ref).max(axis=1), 1e-9)
    assert err.max() < TOL, f'{err.max():.2e} at {err.argmax()}'
