"""Oracle pinning for the SOCP path (BASELINE config 3).  The IPM-CUDA kernel is not built yet; what is
pinned here is the oracle it will be checked against: the unmodified vendored ECOS 2.0.8 (oracle/_ref/libecos_ref.so),
an analytic known answer, the optimality conditions of the hand-derived portfolio form (cvxpygen_b200.families.
portfolio_socp) and the committed golden vectors."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from cvxpygen_b200 import families
from helpers import GOLDEN
from oracle import ref_ecos

needs_ecos = pytest.mark.skipif(not ref_ecos.available(), reason='oracle/_ref/libecos_ref.so not built')


def test_portfolio_family_dimensions_match_survey():
    fam = families.portfolio_socp()
    assert (fam.n_var, fam.n_eq, fam.n_ineq) == (512, 111, 715)
    assert fam.canon_matrix('A').nnz == 917 and fam.canon_matrix('G').nnz == 1314
    assert fam.cone_dims == {'l': 601, 'q': [12, 102]}
    assert fam.changes('c', ['a']) and fam.changes('b', ['w_prev']) and not fam.changes('G')


def _cone_ok(v, l, q, tol):
    ok = (v[:l] >= -tol).all()
    o = l
    for d in q:
        ok &= v[o] >= np.linalg.norm(v[o + 1:o + d]) - tol
        o += d
    return ok


def test_golden_portfolio_solutions_satisfy_optimality_conditions():
    """KKT of  min c'x  s.t. Ax=b, s=h-Gx in K:  A'y + G'z + c = 0, z in K*, s'z = 0."""
    g = np.load(os.path.join(GOLDEN, 'socp_portfolio_100_10.npz'))
    fam = families.portfolio_socp()
    A, G = fam.canon_matrix('A'), fam.canon_matrix('G')
    c0, b0, h = fam.canon_data('c'), fam.canon_data('b'), fam.canon_data('h')
    assert (g['exitflag'] == 0).all()
    for k in range(g['x'].shape[0]):
        c = c0.copy(); c[:100] = -g['param_a'][k]
        b = b0.copy(); b[11:111] = -g['param_w_prev'][k]
        x, y, z, s = g['x'][k], g['y'][k], g['z'][k], g['s'][k]
        assert np.abs(A @ x - b).max() < 1e-7
        assert np.abs(h - G @ x - s).max() < 1e-7
        assert np.abs(A.T @ y + G.T @ z + c).max() < 1e-7
        assert _cone_ok(s, 601, [12, 102], 1e-7) and _cone_ok(z, 601, [12, 102], 1e-7)
        assert abs(s @ z) < 1e-6 and abs(c @ x - g['pcost'][k]) < 1e-7
        w = x[:100]
        assert abs(w.sum() - 1) < 1e-8 and np.abs(w).sum() <= 1.6 + 1e-7


@needs_ecos
def test_ecos_reference_known_answer():
    """min t  s.t. ||(x1 - 1, x2 - 2)|| <= t,  x1 + x2 = 1   ->   x = (0, 1), t = sqrt(2)."""
    c = np.array([0., 0., 1.])
    A = sp.csc_matrix([[1., 1., 0.]]); b = np.array([1.])
    G = sp.csc_matrix(-np.array([[0., 0., 1.], [1., 0., 0.], [0., 1., 0.]])); h = np.array([0., -1., -2.])
    r = ref_ecos.RefECOS(c, A, b, G, h, 0, [3])
    out = r.solve_batch()
    assert out['exitflag'][0] == 0
    assert np.allclose(out['x'][0], [0., 1., np.sqrt(2)], atol=1e-7)


@needs_ecos
def test_ecos_reference_reproduces_golden():
    g = np.load(os.path.join(GOLDEN, 'socp_portfolio_100_10.npz'))
    fam = families.portfolio_socp()
    c0, b0 = fam.canon_data('c'), fam.canon_data('b')
    r = ref_ecos.RefECOS(c0, fam.canon_matrix('A'), b0, fam.canon_matrix('G'), fam.canon_data('h'), 601, [12, 102])
    B = 6
    Cb = np.tile(c0, (B, 1)); Cb[:, :100] = -g['param_a'][:B]
    Bb = np.tile(b0, (B, 1)); Bb[:, 11:111] = -g['param_w_prev'][:B]
    out = r.solve_batch(c=Cb, b=Bb)
    assert np.array_equal(out['iter'], g['iter'][:B]) and np.allclose(out['x'], g['x'][:B], rtol=0, atol=1e-12)


def test_numpy_ipm_restatement_reaches_the_reference_optimum():
    """oracle/ipm_numpy.py (ECOS's algorithm without equilibration, dense KKT solves) against the golden vectors of the
    compiled reference: primal x and equality duals y within the 1e-5 parity bar, identical objective.
    The cone multipliers z of the AUXILIARY epigraph rows are not compared: the hand-derived form is dual degenerate
    there (two interior-point codes at 1e-8 disagree by 1e-3 on them), while the user-level duals are well determined."""
    from oracle.ipm_numpy import ecos_ipm
    g = np.load(os.path.join(GOLDEN, 'socp_portfolio_100_10.npz'))
    fam = families.portfolio_socp()
    c0, b0, h = fam.canon_data('c'), fam.canon_data('b'), fam.canon_data('h')
    A, G = fam.canon_matrix('A'), fam.canon_matrix('G')
    for k in (0, 3):
        c = c0.copy(); c[:100] = -g['param_a'][k]
        b = b0.copy(); b[11:111] = -g['param_w_prev'][k]
        r = ecos_ipm(c, A, b, G, h, 601, [12, 102])
        assert r['exitflag'] == 0 and abs(r['iter'] - g['iter'][k]) <= 3
        assert np.linalg.norm(r['x'] - g['x'][k]) / np.linalg.norm(g['x'][k]) < 1e-5
        assert np.linalg.norm(r['y'] - g['y'][k]) / np.linalg.norm(g['y'][k]) < 1e-5
        assert abs(r['pcost'] - g['pcost'][k]) < 1e-8
        zl1 = fam.duals[2].indices[0]                      # user dual of ||w||_1 <= L
        assert abs(r['z'][zl1] - g['z'][k][zl1]) < 1e-5 * max(1.0, abs(g['z'][k][zl1]))
