import sys, numpy as np
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
from cvxpygen_b200 import standard, families
from helpers import canon_matrix_batches
from oracle.grad_numpy import qp_backward_mat, param_gradient_mat
name, B = 'actuator_1_3', 2000
fam = standard.STANDARD[name][0]()
params = families.actuator_batch(fam, B, seed=2)
mod = standard.load(name, device=0)
res = mod.solve_batch(params, return_canonical=True)
Px, Ax, (q, l, u) = canon_matrix_batches(fam, params, B)
dprim = np.zeros((B, 2)); dprim[:, 0] = 1.0
g, dq, dl, du, dP, dA = mod.gradient_batch_mat(params, res.sol_x, res.sol_y, dprim, return_canonical=True)
n = 64
dx = np.zeros((n, fam.n_var)); dx[:, 0] = 1.0
rq, rl, ru, rP, rA = qp_backward_mat(fam.patterns['P'], fam.patterns['A'], Px[:n], Ax[:n], res.sol_x[:n], res.sol_y[:n], dx)
for nm, a, b in (('dq', dq[:n], rq), ('dl+du', (dl + du)[:n], rl + ru), ('dP', dP[:n], rP), ('dA', dA[:n], rA)):
    d = np.abs(a - b); k = np.unravel_index(d.argmax(), d.shape)
    print(nm, 'maxdiff', d.max(), 'at', k, 'got', a[k], 'want', b[k])
names = standard.STANDARD[name][1]
want = param_gradient_mat(fam, rq, rl, ru, rP, rA, names)
got = np.concatenate([g[nm][:n] for nm in names], axis=1)
d = np.abs(got - want); k = np.unravel_index(d.argmax(), d.shape)
print('dtheta maxdiff', d.max(), 'at', k, 'got', got[k], 'want', want[k])
print('row got ', got[k[0]]); print('row want', want[k[0]]); print('y', res.sol_y[k[0]]); print('x', res.sol_x[k[0]])
