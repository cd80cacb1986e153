#!/usr/bin/env python
"""Diagnostic: solve the fresh 400-instance portfolio batch of tests/test_socp_ipm.py on the GPU twice and save the
canonical solutions (gpurun_out/diag_socp.npz) for offline comparison with the host emulation / the reference."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cvxpygen_b200 import standard, runtime

rng = np.random.default_rng(7)
B = 400
a = rng.standard_normal((B, 100)) * rng.uniform(0.2, 2.0, (B, 1))
wp = np.abs(1 / 100 + 0.02 * rng.standard_normal((B, 100)))
m = runtime.load(sys.argv[1]) if len(sys.argv) > 1 else standard.load('portfolio_socp_100_10')
r1 = m.solve_batch({'a': a, 'w_prev': wp}, return_canonical=True)
r2 = m.solve_batch({'a': a, 'w_prev': wp}, return_canonical=True)
print('bitwise reproducible:', all(np.array_equal(getattr(r1, k), getattr(r2, k)) for k in ('sol_x', 'sol_y', 'sol_z', 'sol_s')),
      'iters equal:', np.array_equal(r1.cpg_info.iter, r2.cpg_info.iter))
os.makedirs('gpurun_out', exist_ok=True)
np.savez('gpurun_out/diag_socp.npz', x=r1.sol_x, y=r1.sol_y, z=r1.sol_z, s=r1.sol_s, it=r1.cpg_info.iter, st=r1.cpg_info.status,
         z2=r2.sol_z)
