"""Pins the oracle (test infrastructure) before anything is checked against it:
  1. known-answer tests of the reference's own suite (OSQP tests/*/generate_problem.py values);
  2. the committed golden vectors generated with the unmodified compiled reference (tests/golden);
  3. when oracle/_ref is present (build container, GPU box): live agreement numpy restatement <-> reference."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from oracle.admm_numpy import AdmmOracle
from cvxpygen_b200 import standard
from helpers import GOLDEN, oracle_for, ref_available, rel_err

INF = np.inf
TIGHT = dict(eps_abs=1e-7, eps_rel=1e-7, max_iter=20000)


def test_kat_basic_qp():
    # osqp_sources/tests/basic_qp/generate_problem.py:5-25
    P = sp.csc_matrix([[4., 1.], [1., 2.]]); q = np.ones(2)
    A = sp.csc_matrix(np.array([[1., 1.], [1., 0.], [0., 1.], [0., 1.]]))
    l = np.array([1., 0., 0., -INF]); u = np.array([1., 0.7, 0.7, INF])
    o = AdmmOracle(P, q, A, l, u, **TIGHT)
    r = o.solve_batch()
    assert r['status'][0] == 1
    assert np.allclose(r['x'][0], [0.3, 0.7], atol=1e-5)
    assert np.allclose(r['y'][0], [-2.9, 0.0, 0.2, 0.0], atol=1e-4)
    assert abs(r['obj'][0] - 1.88) < 1e-5
    # update case of the same KAT (q_new, l_new, u_new): must stay consistent with a fresh setup on the new data
    qn = np.array([2.5, 3.2]); ln = np.array([0.8, -3.4, -INF, 0.5]); un = np.array([1.6, 1.0, INF, 0.5])
    r2 = o.solve_batch(q=qn[None], l=ln[None], u=un[None])
    r3 = AdmmOracle(P, qn, A, ln, un, **TIGHT).solve_batch()
    assert r2['status'][0] == 1 and np.allclose(r2['x'], r3['x'], atol=1e-5)


def test_kat_basic_qp2():
    # osqp_sources/tests/basic_qp2/generate_problem.py:5-36
    P = sp.csc_matrix([[11., 0.], [0., 0.]]); q = np.array([3., 4.])
    A = sp.csc_matrix(np.array([[-1., 0.], [0., -1.], [-1., 3.], [2., 5.], [3., 4.]]))
    l = -INF * np.ones(5); u = np.array([0., 0., -15., 100., 80.])
    o = AdmmOracle(P, q, A, l, u, **TIGHT)
    r = o.solve_batch()
    assert r['status'][0] == 1
    assert np.allclose(r['x'][0], [15., 0.], atol=1e-4)
    assert np.allclose(r['y'][0], [0., 508., 168., 0., 0.], atol=2e-2)
    assert abs(r['obj'][0] - 1282.5) < 1e-2
    r = o.solve_batch(q=np.array([[1., 1.]]), u=np.array([[-2., 0., -20., 100., 80.]]))
    assert np.allclose(r['x'][0], [20., 0.], atol=1e-4) and abs(r['obj'][0] - 2220.0) < 2e-2


def test_kat_primal_dual_infeasibility():
    # osqp_sources/tests/primal_dual_infeasibility/generate_problem.py:5-37 (settings: scaling=0, test header :34-39)
    P = sp.diags([1., 0.], format='csc'); q = np.array([1., -1.])
    A12 = sp.csc_matrix([[1., 1.], [1., 0.], [0., 1.]]); A34 = sp.csc_matrix([[1., 0.], [1., 0.], [0., 1.]])
    l = np.array([0., 1., 1.])
    kw = dict(scaling=0, max_iter=2000)
    r1 = AdmmOracle(P, q, A12, l, np.array([5., 3., 3.]), **kw, eps_abs=1e-7, eps_rel=1e-7).solve_batch()
    assert r1['status'][0] == 1 and np.allclose(r1['x'][0], [1., 3.], atol=1e-4)
    assert np.allclose(r1['y'][0], [0., -2., 1.], atol=1e-3) and abs(r1['obj'][0] + 1.5) < 1e-4
    r2 = AdmmOracle(P, q, A12, l, np.array([0., 3., 3.]), **kw).solve_batch()
    assert r2['status'][0] == -3 and np.isnan(r2['x']).all() and r2['obj'][0] == 1e30
    r3 = AdmmOracle(P, q, A34, l, np.array([2., 3., INF]), **kw).solve_batch()
    assert r3['status'][0] == -4 and r3['obj'][0] == -1e30
    r4 = AdmmOracle(P, q, A34, l, np.array([0., 3., INF]), **kw).solve_batch()
    assert r4['status'][0] == -3


def test_kat_unconstrained():
    # osqp_sources/tests/unconstrained/generate_problem.py:5-16
    P = sp.diags([0.617022, 0.92032449, 0.20011437, 0.50233257, 0.34675589], format='csc')
    q = np.array([-1.10593508, -1.65451545, -2.3634686, 1.13534535, -1.01701414])
    r = AdmmOracle(P, q, sp.csc_matrix((0, 5)), np.zeros(0), np.zeros(0), **TIGHT).solve_batch()
    assert r['status'][0] == 1
    assert np.allclose(r['x'][0], [1.79237542, 1.79775228, 11.81058885, -2.26014678, 2.93293975], atol=1e-4)
    assert abs(r['obj'][0] + 19.209752026813277) < 1e-5


@pytest.mark.parametrize('name', list(standard.QP_NAMES))
@pytest.mark.parametrize('tag,kw', [('default', {}), ('tight', dict(eps_abs=1e-7, eps_rel=1e-7)), ('norho', dict(adaptive_rho=0))])
def test_numpy_oracle_matches_golden(name, tag, kw):
    g = np.load(os.path.join(GOLDEN, f'{name}.npz'))
    fam = standard.STANDARD[name][0]()
    o = oracle_for(fam, **kw)
    if tag == 'default':
        # the restatement sums in numpy's order; the product's equilibration is pinned bit-exactly in test_offline.py
        assert np.allclose(o.D, g['D'], rtol=1e-13) and np.allclose(o.E, g['E'], rtol=1e-13) and abs(o.c / float(g['c']) - 1) < 1e-13
    B = 16                                   # a slice keeps the CPU suite fast
    r = o.solve_batch(q=g['q'][:B], l=g['l'][:B], u=g['u'][:B])
    assert (r['status'] == g[f'{tag}_status'][:B]).all()
    assert (r['iter'] == g[f'{tag}_iter'][:B]).all()
    assert (r['rho_updates'] == g[f'{tag}_rho_updates'][:B]).all()
    ok = np.isin(r['status'], [1, 2, -2])      # no-solution statuses: OSQP 0.6.2 stores (c_float)0x7fc00000UL = 2143289344.0
    assert ok.any()                            # (include/constants.h:96), the restatement and the kernel store an IEEE NaN
    assert rel_err(r['x'][ok], g[f'{tag}_x'][:B][ok]).max() < 1e-9
    assert rel_err(r['y'][ok], g[f'{tag}_y'][:B][ok]).max() < 1e-8
    assert np.allclose(r['obj'][ok], g[f'{tag}_obj'][:B][ok], rtol=1e-9, atol=1e-11)


@pytest.mark.skipif(not ref_available(), reason='oracle/_ref/libosqp_ref.so not built')
def test_reference_library_reproduces_golden():
    from oracle.ref_osqp import RefOSQP
    name = 'mpc_6_3_10'
    g = np.load(os.path.join(GOLDEN, f'{name}.npz'))
    fam = standard.STANDARD[name][0]()
    r = RefOSQP(fam.canon_matrix('P'), fam.canon_data('q'), fam.canon_matrix('A'), fam.canon_data('l'), fam.canon_data('u'))
    s = r.solve_batch(q=g['q'], l=g['l'], u=g['u'])
    assert np.array_equal(s['x'], g['default_x']) and np.array_equal(s['iter'], g['default_iter'])
    assert r.adaptive_rho_interval() == 100       # osqp.c:267-279 without a profiling timer


def test_kat_non_convex():
    """OSQP's non_cvx test (osqp_sources/tests/non_cvx: P indefinite): with the default sigma the SETUP must fail (OSQP_NONCVX_ERROR,
    test_non_cvx.h:35-36 -- here the offline inertia check of the KKT factor); with sigma = 5 setup succeeds and the SOLVE reports
    OSQP_NON_CVX with a NaN objective (:53-58) -- restatement, compiled reference (when built) and the main kernel on the emulator."""
    import scipy.sparse as sp
    from cvxpygen_b200.ir import CanonFamily
    from cvxpygen_b200.offline.qp_setup import setup_qp_family
    from oracle.admm_numpy import AdmmOracle, NON_CVX
    P = sp.triu(sp.csc_matrix([[2., 5.], [5., 1.]]), format='csc'); q = np.array([3., 4.])
    A = sp.csc_matrix([[-1., 0.], [0., -1.], [-1., 3.], [2., 5.], [3., 4]])
    l = -1e30 * np.ones(5); u = np.array([0., 0., -15., 100., 80.])
    fam = CanonFamily.from_canonical_qp('osqp_non_cvx', P, q, A, l, u, n_eq=0)
    with pytest.raises(ValueError, match='non-convex'):
        setup_qp_family(fam, ['q', 'l', 'u'])
    o = AdmmOracle(P, q, A, l, u, sigma=5.0).solve_batch(B=1)
    assert o['status'][0] == NON_CVX and np.isnan(o['obj'][0])
    from helpers import ref_available
    if ref_available():
        from oracle.ref_osqp import RefOSQP
        r = RefOSQP(P, q, A, l, u, sigma=5.0).solve_batch(B=1)
        assert r['status'][0] == NON_CVX and (np.isnan(r['obj'][0]) or r['obj'][0] == 2143289344.0)    # OSQP_NAN as stored by 0.6.2
        assert r['iter'][0] == o['iter'][0]


def test_kat_non_convex_on_the_kernel(tmp_path):
    import scipy.sparse as sp
    from cvxpygen_b200 import codegen
    from cvxpygen_b200.ir import CanonFamily
    from cvxpygen_b200.offline.qp_setup import setup_qp_family
    from oracle.admm_numpy import AdmmOracle, NON_CVX
    import test_simt_emulation as emu
    P = sp.triu(sp.csc_matrix([[2., 5.], [5., 1.]]), format='csc'); q = np.array([3., 4.])
    A = sp.csc_matrix([[-1., 0.], [0., -1.], [-1., 3.], [2., 5.], [3., 4]])
    l = -1e30 * np.ones(5); u = np.array([0., 0., -15., 100., 80.])
    fam = CanonFamily.from_canonical_qp('osqp_non_cvx', P, q, A, l, u, n_eq=0)
    orig = emu.setup_qp_family
    emu.setup_qp_family = lambda f, b: orig(f, b, sigma=5.0)            # the KAT's sigma_new
    try:
        st, lib, dims = emu.build_emu(fam, ['q', 'l', 'u'], str(tmp_path))
    finally:
        emu.setup_qp_family = orig
    rows = np.concatenate([q, l, u])[None, :]
    out = emu.run_solve(lib, 'emu_main_solve', dims, rows, grid=1)
    o = AdmmOracle(P, q, A, l, u, sigma=5.0).solve_batch(B=1)
    assert out['status'][0] == NON_CVX == o['status'][0] and out['iter'][0] == o['iter'][0]
    assert np.isnan(out['obj'][0]) and np.isnan(out['x']).all()
