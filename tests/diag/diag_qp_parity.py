import sys, numpy as np
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
from cvxpygen_b200 import standard
from helpers import family_and_batch, oracle_for, rel_err
for name, kw in [('nonneg_LS_3_2', dict()), ('random_qp_20_5_15', dict(adaptive_rho_interval=25, eps_abs=1e-5, eps_rel=1e-5)), ('mpc_12_4_10', dict(adaptive_rho_interval=25, eps_abs=1e-5, eps_rel=1e-5))]:
    B = 256
    fam, params, (q, l, u) = family_and_batch(name, B, seed=1 if not kw else 5)
    mod = standard.load(name)
    res = mod.solve_batch(params, return_canonical=True, **kw)
    mod.set_solver_default_settings()
    ora = oracle_for(fam, **kw).solve_batch(q=q, l=l, u=u)
    st = res.cpg_info.status
    print(name, 'status mismatch', (st != ora['status']).sum(), 'iter mismatch', (res.cpg_info.iter != ora['iter']).sum(), 'rho_updates', ora['rho_updates'].sum(), (ora['rho_updates']>0).sum())
    bad = np.nonzero((st != ora['status']) | (res.cpg_info.iter != ora['iter']))[0]
    print('  bad idx', bad[:10], 'gpu st/it', st[bad[:10]], res.cpg_info.iter[bad[:10]], 'ora', ora['status'][bad[:10]], ora['iter'][bad[:10]], 'ora rho upd', ora['rho_updates'][bad[:10]])
    sol = np.isin(ora['status'], [1,2,-2]) & np.isin(st, [1,2,-2])
    ex = rel_err(res.sol_x[sol], ora['x'][sol]); ey = rel_err(res.sol_y[sol], ora['y'][sol])
    tail = ora['rho_updates'][sol] > 0
    print('  relerr x max', ex.max(), 'y max', ey.max(), ' tail-only x', ex[tail].max() if tail.any() else None, 'nontail x', ex[~tail].max() if (~tail).any() else None)
    print('  pri_res maxrel', np.max(np.abs(res.cpg_info.pri_res[sol]-ora['pri_res'][sol])/(np.abs(ora['pri_res'][sol])+1e-30)), 'dua', np.max(np.abs(res.cpg_info.dua_res[sol]-ora['dua_res'][sol])/(np.abs(ora['dua_res'][sol])+1e-30)))
    print('  obj maxrel', np.max(np.abs(res.cpg_info.obj_val[sol]-ora['obj'][sol])/(np.abs(ora['obj'][sol])+1e-12)))
