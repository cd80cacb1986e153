"""Small solves of every kernel family, meant to run under compute-sanitizer (memcheck / racecheck):
   compute-sanitizer --tool memcheck python tools/sanitize_run.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
from cvxpygen_b200 import standard, families
from helpers import family_and_batch, ltv_batch

which = sys.argv[1:] or ['mpc', 'ltv', 'big', 'socp', 'socp_mat', 'grad']
if 'mpc' in which:            # main kernel + tail kernel (rho updates every 25 iterations hand instances off), warm start
    fam, params, _ = family_and_batch('mpc_6_3_10', 96, seed=1)
    mod = standard.load('mpc_6_3_10')
    r = mod.solve_batch(params, return_canonical=True, adaptive_rho_interval=25, eps_abs=1e-5, eps_rel=1e-5)
    mod.set_solver_default_settings()
    r2 = mod.solve_batch(params, x0=r.sol_x, y0=r.sol_y, return_canonical=True)
    print('mpc ok', r.cpg_info.iter.mean(), r2.cpg_info.iter.mean())
    if 'grad' in which:
        g = mod.gradient_batch(r.sol_y, np.ones((96, mod.dims.n_prim)), return_canonical=True)
        print('grad ok', float(np.abs(g[1]).max()))
if 'ltv' in which:            # matrix-parameter kernel (equilibration in the factor storage, coloured factorisation)
    fam = standard.STANDARD['mpc_ltv_6_3_10'][0]()
    mod = standard.load('mpc_ltv_6_3_10')
    r = mod.solve_batch(ltv_batch(fam, 64, seed=3), return_canonical=True)
    print('ltv ok', r.cpg_info.iter.mean())
if 'big' in which:            # per-instance-factor kernel with tables read through L2
    fam, params, _ = family_and_batch('random_qp_700_100_700', 8, seed=7)
    r = standard.load('random_qp_700_100_700').solve_batch(params)
    print('big ok', r.cpg_info.iter.mean())
if 'socp' in which:
    g = np.load(os.path.join(ROOT, 'tests', 'golden', 'socp_portfolio_100_10.npz'))
    r = standard.load('portfolio_socp_100_10').solve_batch({'a': g['param_a'][:4], 'w_prev': g['param_w_prev'][:4]})
    print('socp ok', r.cpg_info.iter.mean())
if 'socp_mat' in which:
    fc = families.portfolio_socp(20, 4, matrix_params=True)
    rng = np.random.default_rng(3)
    pc = dict(a=fc.param('a').default[None, :] + 0.3 * rng.standard_normal((8, 20)), w_prev=np.full((8, 20), 1 / 20),
              F=fc.param('F').default[None, :] + 0.25 * rng.standard_normal((8, 80)),
              d_sqrt=fc.param('d_sqrt').default[None, :] * rng.uniform(0.5, 1.5, (8, 20)))
    r = standard.load('portfolio_socp_mat_20_4').solve_batch(pc)
    print('socp_mat ok', r.cpg_info.iter.mean())
