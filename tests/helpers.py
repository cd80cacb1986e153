"""Shared test helpers: canonical batches for a family and the oracle solve on them.

Oracle choice: `oracle_solve` uses the compiled unmodified reference (oracle/_ref/libosqp_ref.so) when it is
present -- it is ~100x faster than the numpy restatement, which matters on the GPU box where test time is
GPU budget -- and the numpy restatement (oracle/admm_numpy.py) otherwise.  tests/test_oracle.py pins the two
against each other and against the golden vectors."""
import os

import numpy as np

from cvxpygen_b200 import families, standard

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def family_and_batch(name, B, seed=1):
    """Returns (family, params dict for the batched parameters, q/l/u canonical batches (unscaled))."""
    fam = standard.STANDARD[name][0]()
    batch = standard.STANDARD[name][1]
    rng = np.random.default_rng(seed)
    params = {}
    for pn in batch:
        p = fam.param(pn)
        if name.startswith('mpc'):
            params[pn] = rng.uniform(-1, 1, (B, p.size))
        else:
            params[pn] = np.asarray(p.default)[None, :] + 0.3 * rng.standard_normal((B, p.size))
    if name in standard.BIG_NAMES:        # keep the 1 500-row family feasible: equalities barely moved, inequalities only loosened
        params['b'] = np.asarray(fam.param('b').default)[None, :] + 0.002 * rng.standard_normal((B, fam.param('b').size))
        params['h'] = np.asarray(fam.param('h').default)[None, :] + 0.3 * np.abs(rng.standard_normal((B, fam.param('h').size)))
    if name.startswith('box_qp'):
        params['q'][:, -1] = 0.0          # keep the curvature-less variable free of cost: bounded instances
    return fam, params, canon_batches(fam, params, B)


def canon_batches(fam, params, B):
    th = np.tile(fam.theta_default(), (B, 1))
    for pn, v in params.items():
        p = fam.param(pn)
        th[:, p.col:p.col + p.size] = v
    q = np.asarray(th @ fam.maps['q'].T.toarray())
    l = np.clip(np.asarray(th @ fam.maps['l'].T.toarray()), -1e30, 1e30)
    u = np.clip(np.asarray(th @ fam.maps['u'].T.toarray()), -1e30, 1e30)
    return q, l, u


def oracle_for(fam, **settings):
    from oracle.admm_numpy import AdmmOracle
    return AdmmOracle(fam.canon_matrix('P'), fam.canon_data('q'), fam.canon_matrix('A'),
                      fam.canon_data('l'), fam.canon_data('u'), **settings)


def ref_available():
    from oracle import ref_osqp
    return ref_osqp.available()


def oracle_solve(fam, q, l, u, x0=None, y0=None, prefer_ref=True, **settings):
    """dict(x, y, obj, iter, status, pri_res, dua_res, rho_updates) for the canonical batch."""
    if prefer_ref and ref_available():
        from oracle.ref_osqp import RefOSQP
        r = RefOSQP(fam.canon_matrix('P'), fam.canon_data('q'), fam.canon_matrix('A'),
                    fam.canon_data('l'), fam.canon_data('u'), nthreads=min(8, os.cpu_count() or 1), **settings)
        return r.solve_batch(q=q, l=l, u=u, x0=x0, y0=y0)
    return oracle_for(fam, **settings).solve_batch(q=q, l=l, u=u, x0=x0, y0=y0)


def rel_err(a, b):
    nb = np.linalg.norm(b, axis=1)
    return np.linalg.norm(a - b, axis=1) / np.maximum(nb, 1e-12)


def assert_batch_parity(got_x, got_y, got_info, ora, tol, eps_abs=1e-3, stable=None, obj_sign=1.0):
    """got_info: object with status, iter, obj_val, pri_res, dua_res arrays.
    stable: optional boolean mask of the instances on which the comparison is meaningful (see rounding_stable)."""
    st = np.asarray(got_info.status)
    assert (st != -100).all(), 'instances left in the hand-off state: the tail kernel did not run'
    if stable is None:
        stable = np.ones(len(st), bool)
    bad = stable & (st != ora['status'])
    assert not bad.any(), f"status mismatch at {np.nonzero(bad)[0][:5]}"
    bad = stable & (np.asarray(got_info.iter) != ora['iter'])
    assert not bad.any(), f"iter mismatch at {np.nonzero(bad)[0][:5]}"
    sol = np.isin(st, [1, 2, -2]) & stable
    if sol.any():
        assert rel_err(got_x[sol], ora['x'][sol]).max() < tol
        assert rel_err(got_y[sol], ora['y'][sol]).max() < tol
        # cpg_retrieve_info reports -(solver objective) for maximisation problems (cvxpygen/utils.py:980)
        assert np.allclose(got_info.obj_val[sol], obj_sign * ora['obj'][sol], rtol=1e-6, atol=1e-9)
        # residuals are differences of O(1) numbers: compare on the scale of the stopping tolerance
        assert np.allclose(got_info.pri_res[sol], ora['pri_res'][sol], rtol=1e-5, atol=1e-3 * eps_abs)
        assert np.allclose(got_info.dua_res[sol], ora['dua_res'][sol], rtol=1e-5, atol=1e-3 * eps_abs)
    nos = ~np.isin(st, [1, 2, -2]) & stable
    if nos.any():
        assert np.isnan(got_x[nos]).all() and np.isnan(got_y[nos]).all()
    return sol


def rounding_stable(fam, q, l, u, ora, **settings):
    """Instances on which the compiled reference and the numpy restatement -- the same algorithm with a different, equally
    exact KKT solve -- stop at the same iteration.  A handful of ill-conditioned instances (e.g. loose rows, rho = 1e-6,
    several rho updates, hundreds of iterations) amplify rounding differences into different stopping iterations; for
    those "the reference's result" is itself not defined to 1e-5 and they are excluded from the comparison."""
    o2 = oracle_for(fam, **settings).solve_batch(q=q, l=l, u=u)
    return (o2['iter'] == ora['iter']) & (o2['status'] == ora['status'])


def conic_batch(fam, B, seed):
    """Parameter batch for the generic conic families (cvxpygen_b200.families.random_socp): perturbed defaults, with one
    instance in eight primal infeasible (LP rows 0/1 contradict) and one in eight unbounded (free direction with negative cost).
    Returns (params dict, kind) with kind 0 = nominal, 1 = primal infeasible, 2 = dual infeasible."""
    rng = np.random.default_rng(seed)
    n, p, m = fam.n_var, fam.n_eq, fam.n_ineq
    kind = rng.integers(0, 8, B)
    kind = np.where(kind < 6, 0, kind - 5)
    c = fam.param('c').default + 0.1 * rng.standard_normal((B, n))
    h = fam.param('h').default + 0.05 * np.abs(rng.standard_normal((B, m)))
    out = {'c': c, 'h': h}
    if p:
        out['b'] = fam.param('b').default + 0.02 * rng.standard_normal((B, p))
    h[kind == 1, 1] = -h[kind == 1, 0] - 1.0
    c[kind == 2, n - 1] = -1.0
    return out, kind


# ---------------------------------------------------------------------------------------------------------
# families with per-instance MATRIX parameters
def ltv_batch(fam, B, seed=3, spread=0.05):
    return families.mpc_ltv_batch(fam, B, seed=seed, spread=spread)


def canon_matrix_batches(fam, params, B):
    """(Px (B, nnzP), Ax (B, nnzA)) canonical matrix entries in CSC order + (q, l, u) batches."""
    th = np.tile(fam.theta_default(), (B, 1))
    for pn, v in params.items():
        p = fam.param(pn)
        th[:, p.col:p.col + p.size] = v
    Px = np.asarray((fam.maps['P'] @ th.T).T)
    Ax = np.asarray((fam.maps['A'] @ th.T).T)
    return Px, Ax, canon_batches(fam, params, B)


def matrix_oracle_solve(fam, Px, Ax, q, l, u, prefer_ref=True, **settings):
    """Reference result for per-instance matrices: osqp_update_P_A + vector updates + solve from the pristine workspace
    (compiled OSQP 0.6.2 when present, else the numpy restatement)."""
    import scipy.sparse as sp
    if prefer_ref and ref_available():
        from oracle.ref_osqp import RefOSQP
        r = RefOSQP(fam.canon_matrix('P'), fam.canon_data('q'), fam.canon_matrix('A'),
                    fam.canon_data('l'), fam.canon_data('u'), nthreads=min(8, os.cpu_count() or 1), **settings)
        assert (Px is None or r._keep[0].nnz == Px.shape[1]) and (Ax is None or r._keep[1].nnz == Ax.shape[1]), \
            'structural zeros were dropped'
        return r.solve_batch_mat(Px=Px, Ax=Ax, q=q, l=l, u=u)
    from oracle.admm_numpy import solve_matrix_batch
    from cvxpygen_b200.offline.qp_setup import unscale_roundtrip
    from cvxpygen_b200.offline.equilibrate import ruiz_equilibrate
    Pi, Pp, Ps = fam.patterns['P']; Ai, Ap, As = fam.patterns['A']
    sc = ruiz_equilibrate(fam.canon_matrix('P'), fam.canon_matrix('A'), fam.canon_data('q'), int(settings.get('scaling', 10)))
    q_un = unscale_roundtrip(sc, int(settings.get('scaling', 10)))[2]
    B = (Px if Px is not None else Ax).shape[0]
    Pl = [sp.csc_matrix((Px[i], Pi, Pp), shape=Ps) if Px is not None else fam.canon_matrix('P') for i in range(B)]
    Al = [sp.csc_matrix((Ax[i], Ai, Ap), shape=As) if Ax is not None else fam.canon_matrix('A') for i in range(B)]
    return solve_matrix_batch(Pl, Al, q_un, fam.canon_data('l'), fam.canon_data('u'), q=q, l=l, u=u, **settings)
