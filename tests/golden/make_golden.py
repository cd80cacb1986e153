#!/usr/bin/env python
"""Generates tests/golden/*.npz with the UNMODIFIED reference compiled into oracle/_ref/libosqp_ref.so
(vendored OSQP 0.6.2, see oracle/Makefile).  Run in the build container, where /root/reference exists:

    make -C oracle ref && python tests/golden/make_golden.py

Each file holds, for one standard family: the seeded batched parameters, the canonical q/l/u batches, the
reference's scaling (D, E, c) and its solutions with cvxpygen's default OSQP settings (adaptive rho on) and
with a tight tolerance.  The fixtures let the CPU test-suite pin oracle/admm_numpy.py and the offline
pipeline without the reference tree, and the GPU suite pin the kernel on a box where it does not exist."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from helpers import family_and_batch          # noqa: E402
from oracle.ref_osqp import RefOSQP           # noqa: E402
from cvxpygen_b200 import standard            # noqa: E402

B = 48
for name in standard.QP_NAMES:
    fam, params, (q, l, u) = family_and_batch(name, B, seed=2024)
    out = {'B': B, 'seed': 2024}
    for k, v in params.items():
        out['param_' + k] = v
    out.update(q=q, l=l, u=u)
    for tag, kw in (('default', {}), ('tight', dict(eps_abs=1e-7, eps_rel=1e-7)), ('norho', dict(adaptive_rho=0))):
        r = RefOSQP(fam.canon_matrix('P'), fam.canon_data('q'), fam.canon_matrix('A'),
                    fam.canon_data('l'), fam.canon_data('u'), **kw)
        if tag == 'default':
            D, E, c = r.scaling()
            out.update(D=D, E=E, c=c)
        s = r.solve_batch(q=q, l=l, u=u)
        for k in ('x', 'y', 'obj', 'iter', 'status', 'pri_res', 'dua_res', 'rho_updates'):
            out[f'{tag}_{k}'] = s[k]
    np.savez_compressed(os.path.join(HERE, f'{name}.npz'), **out)
    print(name, {t: int(out[f'{t}_iter'].mean()) for t in ('default', 'tight', 'norho')}, 'rho updates', int(out['default_rho_updates'].sum()))
