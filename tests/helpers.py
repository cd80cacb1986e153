"""Shared test helpers: canonical batches for a family and the oracle solve on them."""
import numpy as np

from cvxpygen_b200 import families, standard


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
    return fam, params, canon_batches(fam, params, B)


def canon_batches(fam, params, B):
    th = np.tile(fam.theta_default(), (B, 1))
    for pn, v in params.items():
        p = fam.param(pn)
        th[:, p.col:p.col + p.size] = v
    q = th @ fam.maps['q'].T.toarray() if fam.maps['q'].nnz else np.zeros((B, fam.n_var))
    l = np.clip(np.asarray(th @ fam.maps['l'].T.toarray()), -1e30, 1e30)
    u = np.clip(np.asarray(th @ fam.maps['u'].T.toarray()), -1e30, 1e30)
    return np.asarray(q), l, u


def oracle_for(fam, **settings):
    from oracle.admm_numpy import AdmmOracle
    return AdmmOracle(fam.canon_matrix('P'), fam.canon_data('q'), fam.canon_matrix('A'),
                      fam.canon_data('l'), fam.canon_data('u'), **settings)


def rel_err(a, b):
    nb = np.linalg.norm(b, axis=1)
    return np.linalg.norm(a - b, axis=1) / np.maximum(nb, 1e-12)
