"""Two-stage gradient: a QP family solved by the CONIC backend and differentiated through its QP form, the
reference's route for `gradient=True` with a conic solver (cvxpygen/canonicalizer.py:54-65, 334-406; cvxpygen/writer.py:177-206).

CPU: the conic restatement (stage 2) of two QP families, solved by the compiled reference ECOS, brought back to the QP form by
`intermediate_solution`, IS the QP solution (the compiled OSQP at 1e-10 is the yardstick; the epigraph form limits x to ~sqrt of
the conic tolerance, the reference's own tests allow 10 % for this route: tests/test_diff.py).
GPU: IPM-CUDA forward on the conic family == compiled ECOS; backward kernel of the QP library at the intermediate solution ==
the numpy restatement of cpg_osqp_gradient at the same point; the parameter gradient agrees with the QP-route gradient."""
import os

import numpy as np
import pytest

from cvxpygen_b200 import families, standard
from cvxpygen_b200.two_stage import conic_family_of_qp, intermediate_solution
from helpers import canon_batches, oracle_solve
from oracle import ref_ecos
from oracle.grad_numpy import qp_backward, param_gradient


def _conic_rows(cf, fam, params, B):
    th = np.tile(cf.theta_default(), (B, 1))
    for k, v in params.items():
        p = fam.param(k); th[:, p.col:p.col + p.size] = v
    return np.asarray(th @ cf.maps['c'].T.toarray()), np.asarray(th @ cf.maps['h'].T.toarray())


@pytest.mark.skipif(not ref_ecos.available(), reason='oracle/_ref/libecos_ref.so not built')
@pytest.mark.parametrize('builder,batch', [(lambda: families.mpc(6, 3, 10), ['x_init']), (lambda: families.portfolio_qp(50, 10), ['a', 'w_prev'])])
def test_conic_restatement_of_a_qp_family_solves_the_qp(builder, batch):
    fam = builder()
    cf = conic_family_of_qp(fam, batch)
    assert cf.solver_type == 'conic' and cf.n_eq == 0 and cf.n_var == fam.n_var + 1 and len(cf.cone_dims['q']) == 1
    B = 5
    rng = np.random.default_rng(0)
    params = {k: (rng.uniform(-1, 1, (B, fam.param(k).size)) if k == 'x_init' else
                  np.asarray(fam.param(k).default)[None, :] + 0.3 * rng.standard_normal((B, fam.param(k).size))) for k in batch}
    q, l, u = canon_batches(fam, params, B)
    ora = oracle_solve(fam, q, l, u, eps_abs=1e-10, eps_rel=1e-10, max_iter=200000)
    c, h = _conic_rows(cf, fam, params, B)
    R = ref_ecos.RefECOS(cf.canon_data('c'), cf.canon_matrix('A'), np.zeros(0), cf.canon_matrix('G'), cf.canon_data('h'),
                         cf.cone_dims['l'], cf.cone_dims['q'])
    out = R.solve_batch(c=c, h=h)
    assert (out['exitflag'] == 0).all()
    x, y = intermediate_solution(cf, out['x'], out['z'])
    assert np.allclose(out['pcost'], ora['obj'], rtol=1e-6, atol=1e-7)                 # same optimal value
    assert np.abs(x - ora['x']).max() < 2e-3 * np.abs(ora['x']).max()                   # x: sqrt(conic tolerance)
    assert np.abs(y - ora['y']).max() < 5e-2 * np.abs(ora['y']).max()
    with pytest.raises(ValueError, match='extended DPP'):
        conic_family_of_qp(families.mpc_ltv(4, 2, 5), ['qdiag'])                        # a parameter in P: canonicalizer.py:339-343


@pytest.mark.gpu
def test_two_stage_forward_backward_on_the_gpu():
    name = 'mpc_6_3_10_two_stage'
    fam = standard.STANDARD[name][0]()
    ts = standard.load(name)
    cf = conic_family_of_qp(fam, ['x_init'])
    B = 64
    xi = np.random.default_rng(4).uniform(-1, 1, (B, 6))
    sol = ts.solve_batch({'x_init': xi})
    assert (sol.cpg_info.status == 0).all()
    # forward: the conic solve equals the compiled reference ECOS on the same conic data
    c, h = _conic_rows(cf, fam, {'x_init': xi}, B)
    R = ref_ecos.RefECOS(cf.canon_data('c'), cf.canon_matrix('A'), np.zeros(0), cf.canon_matrix('G'), cf.canon_data('h'),
                         cf.cone_dims['l'], cf.cone_dims['q'])
    ref = R.solve_batch(c=c, h=h)
    assert np.array_equal(sol.cpg_info.iter, ref['iter'])
    assert np.abs(sol.conic.sol_x - ref['x']).max() < 1e-6 * np.abs(ref['x']).max()
    xr, yr = intermediate_solution(cf, ref['x'], ref['z'])
    assert np.abs(sol.sol_x - xr).max() < 1e-6 and np.abs(sol.sol_y - yr).max() < 1e-5
    # user-level variables of the QP family come straight out of the conic x block
    prim_idx = np.concatenate([v.indices for v in fam.variables])
    assert np.array_equal(sol.prim, sol.conic.sol_x[:, prim_idx])
    # backward: the QP library's kernel at the intermediate solution == numpy restatement of cpg_osqp_gradient at that point
    dprim = np.random.default_rng(5).standard_normal((B, len(prim_idx)))
    got = ts.gradient_batch(sol, dprim)['x_init']
    dx = np.zeros((B, fam.n_var)); dx[:, prim_idx] = dprim
    dq, dl, du, _ = qp_backward(fam.canon_matrix('P'), fam.canon_matrix('A'), sol.sol_x, sol.sol_y, dx)
    want = param_gradient(fam, dq, dl, du, ['x_init'])
    assert np.abs(got - want).max() < 1e-5 * np.abs(want).max()
