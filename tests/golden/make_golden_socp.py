#!/usr/bin/env python
"""Golden vectors for BASELINE config 3 (portfolio SOCP, n=100 assets) produced by the UNMODIFIED vendored ECOS 2.0.8
(oracle/_ref/libecos_ref.so, built by `make -C oracle ref`) with cvxpygen's settings (feastol=abstol=reltol=1e-8).
These fixtures pin the IPM-CUDA kernel and its host emulation; the generic conic families add exit flags 1 and 2."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from cvxpygen_b200 import families          # noqa: E402
from oracle.ref_ecos import RefECOS         # noqa: E402

fam = families.portfolio_socp()
c, b, h = fam.canon_data('c'), fam.canon_data('b'), fam.canon_data('h')
A, G = fam.canon_matrix('A'), fam.canon_matrix('G')
r = RefECOS(c, A, b, G, h, fam.cone_dims['l'], fam.cone_dims['q'])
B = 24
rng = np.random.default_rng(2024)
a = rng.standard_normal((B, 100)); wp = 1 / 100 + 0.01 * rng.standard_normal((B, 100))
Cb = np.tile(c, (B, 1)); Cb[:, :100] = -a
Bb = np.tile(b, (B, 1)); Bb[:, 11:111] = -wp
out = r.solve_batch(c=Cb, b=Bb)
np.savez_compressed(os.path.join(HERE, 'socp_portfolio_100_10.npz'), param_a=a, param_w_prev=wp,
                    x=out['x'], y=out['y'], z=out['z'], s=out['s'], pcost=out['pcost'], iter=out['iter'], exitflag=out['exitflag'])
print('iters', out['iter'].mean(), 'exit', np.unique(out['exitflag']), 'us/solve', out['seconds'] / B * 1e6)

# ---- generic conic families (three cones + equalities; pure LP; no equalities): exit flags 0 / 1 / 2
sys.path.insert(0, os.path.dirname(HERE))
from helpers import conic_batch             # noqa: E402
for fam in (families.random_socp(30, 8, 20, (3, 5, 4), seed=5), families.random_socp(20, 5, 30, (), seed=6),
            families.random_socp(12, 0, 10, (6,), seed=7)):
    par, kind = conic_batch(fam, 64, seed=99)
    r = RefECOS(fam.canon_data('c'), fam.canon_matrix('A'), fam.canon_data('b'), fam.canon_matrix('G'), fam.canon_data('h'),
                fam.cone_dims['l'], fam.cone_dims['q'])
    out = r.solve_batch(c=par['c'], h=par['h'], b=par.get('b'))
    np.savez_compressed(os.path.join(HERE, f'socp_{fam.name}.npz'), kind=kind, **{'param_' + k: v for k, v in par.items()},
                        x=out['x'], y=out['y'], z=out['z'], s=out['s'], pcost=out['pcost'], iter=out['iter'], exitflag=out['exitflag'])
    print(fam.name, 'kinds', np.bincount(kind), 'flags', np.unique(out['exitflag'], return_counts=True), 'iters', out['iter'].mean())

# ---- the reference's own LP test problem run with ECOS (tests/test_E2E_LP.py:15-36, 66-74): network flow, n = 50, m = 10
fam = families.network_lp(50, 10)
par = families.network_lp_batch(fam, 48, seed=2025)
th = np.tile(fam.theta_default(), (48, 1))
for k, v in par.items():
    p = fam.param(k); th[:, p.col:p.col + p.size] = v
Cb = np.asarray(th @ fam.maps['c'].T.toarray()); Hb = np.asarray(th @ fam.maps['h'].T.toarray())
r = RefECOS(fam.canon_data('c'), fam.canon_matrix('A'), fam.canon_data('b'), fam.canon_matrix('G'), fam.canon_data('h'),
            fam.cone_dims['l'], fam.cone_dims['q'])
out = r.solve_batch(c=Cb, h=Hb)
np.savez_compressed(os.path.join(HERE, f'socp_{fam.name}.npz'), **{'param_' + k: v for k, v in par.items()},
                    x=out['x'], y=out['y'], z=out['z'], s=out['s'], pcost=out['pcost'], iter=out['iter'], exitflag=out['exitflag'])
print(fam.name, 'flags', np.unique(out['exitflag'], return_counts=True), 'iters', out['iter'].mean())

# ---- the reference's SOCP test problem (tests/test_E2E_SOCP.py:15-63, ADP step): four second-order cones, no LP cone, no equalities
fam = families.adp_socp()
rng = np.random.default_rng(2026)
fb = fam.param('f').default[None, :] + 0.5 * rng.standard_normal((48, 6))
th = np.tile(fam.theta_default(), (48, 1)); p = fam.param('f'); th[:, p.col:p.col + p.size] = fb
Hb = np.asarray(th @ fam.maps['h'].T.toarray())
r = RefECOS(fam.canon_data('c'), fam.canon_matrix('A'), fam.canon_data('b'), fam.canon_matrix('G'), fam.canon_data('h'),
            fam.cone_dims['l'], fam.cone_dims['q'])
out = r.solve_batch(h=Hb)
np.savez_compressed(os.path.join(HERE, f'socp_{fam.name}.npz'), param_f=fb,
                    x=out['x'], y=out['y'], z=out['z'], s=out['s'], pcost=out['pcost'], iter=out['iter'], exitflag=out['exitflag'])
print(fam.name, 'flags', np.unique(out['exitflag'], return_counts=True), 'iters', out['iter'].mean())
