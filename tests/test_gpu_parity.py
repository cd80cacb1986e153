"""GPU parity tests: the sm_100a kernels (through the C ABI, libcpg_b200.so) against the oracle on
identical seeded parameter batches.  Tolerance: 1e-5 relative on primal/dual (BASELINE.json north_star);
iteration counts and statuses must be identical."""
import numpy as np
import pytest

from cvxpygen_b200 import standard
from helpers import family_and_batch, oracle_solve, rel_err, assert_batch_parity

TOL = 1e-5   # north_star: "within 1e-5 relative on primal/dual variables"

FAMS = [('mpc_12_4_10', 512), ('mpc_6_3_10', 512), ('nonneg_LS_3_2', 256), ('random_qp_20_5_15', 256), ('portfolio_qp_50_10', 256)]


@pytest.mark.gpu
@pytest.mark.parametrize('name,B', FAMS)
@pytest.mark.parametrize('adaptive_rho', [0, 1])
def test_parity_vs_oracle(name, B, adaptive_rho):
    fam, params, (q, l, u) = family_and_batch(name, B)
    mod = standard.load(name)
    res = mod.solve_batch(params, return_canonical=True, adaptive_rho=adaptive_rho)
    mod.set_solver_default_settings()
    ora = oracle_solve(fam, q, l, u, adaptive_rho=adaptive_rho)
    sol = assert_batch_parity(res.sol_x, res.sol_y, res.cpg_info, ora, TOL, obj_sign=-1.0 if fam.is_maximization else 1.0)
    # user-level retrieval = gather of the canonical solution (a12)
    for v in fam.variables:
        got = res.cpg_prim[v.name].reshape(B, -1, order='F') if len(v.shape) > 1 else res.cpg_prim[v.name]
        assert np.array_equal(got[sol].reshape(sol.sum(), -1), res.sol_x[sol][:, v.indices])
    for d in fam.duals:
        assert np.array_equal(res.cpg_dual[d.name][sol], res.sol_y[sol][:, d.indices])


@pytest.mark.gpu
@pytest.mark.parametrize('name', ['mpc_12_4_10', 'random_qp_20_5_15', 'nonneg_LS_3_2'])
def test_parity_many_rho_updates(name):
    """adaptive_rho_interval=25 makes a large share of the instances re-factor (several times): exercises the
    tail kernel's numeric LDL' and per-instance triangular solves."""
    B = 384
    fam, params, (q, l, u) = family_and_batch(name, B, seed=5)
    mod = standard.load(name)
    kw = dict(adaptive_rho_interval=25, eps_abs=1e-5, eps_rel=1e-5)
    res = mod.solve_batch(params, return_canonical=True, **kw)
    mod.set_solver_default_settings()
    ora = oracle_solve(fam, q, l, u, **kw)
    assert ora['rho_updates'].sum() > B // 8
    assert_batch_parity(res.sol_x, res.sol_y, res.cpg_info, ora, TOL, eps_abs=1e-5)


@pytest.mark.gpu
def test_warm_start_and_device_api():
    """x0/y0 warm start (osqp_warm_start, osqp.c:929-953) and the device-pointer entry point on torch tensors."""
    import torch
    name, B = 'mpc_12_4_10', 256
    fam, params, (q, l, u) = family_and_batch(name, B, seed=3)
    mod = standard.load(name)
    cold = mod.solve_batch(params, return_canonical=True)
    # perturb the parameters slightly and warm start from the previous solution
    params2 = {'x_init': params['x_init'] + 0.01}
    _, _, (q2, l2, u2) = (None, None, __import__('helpers').canon_batches(fam, params2, B))
    warm = mod.solve_batch(params2, x0=cold.sol_x, y0=cold.sol_y, return_canonical=True)
    ora = oracle_solve(fam, q2, l2, u2, x0=cold.sol_x, y0=cold.sol_y)
    assert_batch_parity(warm.sol_x, warm.sol_y, warm.cpg_info, ora, TOL)
    assert warm.cpg_info.iter.mean() <= cold.cpg_info.iter.mean()
    # device API: same numbers as the host API
    P = torch.from_numpy(mod.pack_params(params2)).cuda()
    x0 = torch.from_numpy(cold.sol_x).cuda(); y0 = torch.from_numpy(cold.sol_y).cuda()
    out = mod.solve_batch_device(P, x0=x0, y0=y0, return_canonical=True)
    torch.cuda.synchronize()
    assert np.array_equal(out.sol_x.cpu().numpy(), warm.sol_x)
    assert np.array_equal(out.status.cpu().numpy(), warm.cpg_info.status)
    assert mod.launch_count() == 2


@pytest.mark.gpu
def test_settings_max_iter_and_inaccurate():
    """max_iter smaller than convergence: OSQP re-checks with 10x looser tolerances (osqp.c:563-568)."""
    name, B = 'mpc_12_4_10', 128
    fam, params, (q, l, u) = family_and_batch(name, B, seed=11)
    mod = standard.load(name)
    for kw in (dict(max_iter=30), dict(max_iter=10, check_termination=0), dict(eps_abs=1e-6, eps_rel=1e-6, max_iter=60)):
        res = mod.solve_batch(params, return_canonical=True, **kw)
        mod.set_solver_default_settings()
        ora = oracle_solve(fam, q, l, u, **kw)
        assert_batch_parity(res.sol_x, res.sol_y, res.cpg_info, ora, TOL, eps_abs=kw.get('eps_abs', 1e-3))
        assert set(np.unique(ora['status'])) <= {1, 2, -2}


@pytest.mark.gpu
def test_single_instance_cpg_solve_api():
    """Reference-style single-instance entry: cpg_module.solve(upd, par) (cvxpygen/utils.py:1194-1270)."""
    name = 'nonneg_LS_3_2'
    mod = standard.load(name)
    fam, params, (q, l, u) = family_and_batch(name, 1, seed=2)
    upd = mod.cpg_updated(); par = mod.cpg_params()
    upd.b = True; par.b = list(params['b'][0])
    res = mod.solve(upd, par)
    ora = oracle_solve(fam, q, l, u)
    assert res.cpg_info.status == 'solved' and res.cpg_info.iter == int(ora['iter'][0])
    assert np.allclose(res.cpg_prim.x, ora['x'][0, :2], rtol=1e-6, atol=1e-9)
    assert abs(res.cpg_info.obj_val - ora['obj'][0]) < 1e-9
    with pytest.raises(AttributeError):
        mod.set_solver_setting('no_such_setting', 1)


@pytest.mark.gpu
def test_full_size_properties():
    """BASELINE config 2 at full size (batch = 100 000): size-independent properties instead of an oracle run.
    (1) every instance reports `solved`; (2) the returned point satisfies OSQP's own stopping test when the
    residuals are recomputed on the host from the returned x, y (unscaled data); (3) permutation equivariance:
    solving a shuffled batch returns the shuffled solutions bit for bit (instances do not interact)."""
    name, B = 'mpc_12_4_10', 100000
    fam, params, (q, l, u) = family_and_batch(name, B, seed=1)
    mod = standard.load(name)
    res = mod.solve_batch(params, return_canonical=True)
    assert (res.cpg_info.status == 1).all()
    P = fam.canon_matrix('P'); P = (P + __import__('scipy.sparse').sparse.triu(P, 1).T).tocsr(); A = fam.canon_matrix('A').tocsr()
    x, y = res.sol_x, res.sol_y
    Ax = (A @ x.T).T
    z = np.clip(Ax, l, u)                                  # closest feasible z
    pri = np.abs(Ax - z).max(axis=1)
    dua = np.abs((P @ x.T).T + q + (A.T @ y.T).T).max(axis=1)
    eps_p = 1e-3 + 1e-3 * np.maximum(np.abs(Ax).max(axis=1), np.abs(z).max(axis=1))
    eps_d = 1e-3 + 1e-3 * np.maximum(np.abs((P @ x.T).T).max(axis=1), np.abs((A.T @ y.T).T).max(axis=1))
    assert (pri <= eps_p).all() and (dua <= 1.05 * eps_d).all()
    perm = np.random.default_rng(0).permutation(B)
    res2 = mod.solve_batch({'x_init': params['x_init'][perm]}, return_canonical=True)
    assert np.array_equal(res2.sol_x, res.sol_x[perm]) or np.allclose(res2.sol_x, res.sol_x[perm], rtol=0, atol=1e-9)
    assert np.array_equal(res2.cpg_info.iter, res.cpg_info.iter[perm])


def corner_case_batch(fam, B, seed=13):
    """box_qp instances: regular, equality-collapsed rows, loose rows, primal infeasible, dual infeasible."""
    rng = np.random.default_rng(seed)
    pq, pl, pu = (fam.param(k) for k in ('q', 'l', 'u'))
    q = np.asarray(pq.default)[None] + 0.3 * rng.standard_normal((B, pq.size)); q[:, -1] = 0.0
    mid = 0.5 * (np.asarray(pl.default) + np.asarray(pu.default))[None] + 0.2 * rng.standard_normal((B, pl.size))
    half = 0.3 + rng.random((B, pl.size))
    l, u = mid - half, mid + half
    kind = np.arange(B) % 5
    for b in range(B):
        if kind[b] == 1:                       # some rows become equalities (u - l < 1e-4)
            rows = rng.choice(np.arange(2, pl.size), 2, replace=False); l[b, rows] = u[b, rows] = mid[b, rows]
        elif kind[b] == 2:                     # some rows become loose
            rows = rng.choice(np.arange(2, pl.size), 2, replace=False); l[b, rows] = -1e30; u[b, rows] = 1e30
        elif kind[b] == 3:                     # parallel rows 0, 1 with crossing bounds: primal infeasible
            l[b, 0] = 2.0; u[b, 0] = 3.0; l[b, 1] = -1.0; u[b, 1] = 1.0
        elif kind[b] == 4:                     # cost on the free, curvature-less variable: dual infeasible
            q[b, -1] = 1.0 + rng.random()
    return {'q': q, 'l': l, 'u': u}, kind


@pytest.mark.gpu
def test_corner_cases_type_changes_and_infeasibility():
    """Per-instance constraint-type changes (routed to the tail kernel with their own KKT factor), primal and dual
    infeasibility certificates (NaN solution, +-1e30 objective): statuses, iteration counts and solutions vs the oracle."""
    from helpers import canon_batches
    name, B = 'box_qp_6_8', 320
    fam = standard.STANDARD[name][0]()
    params, kind = corner_case_batch(fam, B)
    q, l, u = canon_batches(fam, params, B)
    mod = standard.load(name)
    res = mod.solve_batch(params, return_canonical=True)
    ora = oracle_solve(fam, q, l, u)
    assert set(np.unique(ora['status'])) >= {1, -3, -4}
    assert (ora['status'][kind == 3] == -3).all() and (ora['status'][kind == 4] == -4).all()
    from helpers import rounding_stable
    stable = rounding_stable(fam, q, l, u, ora)
    assert stable.mean() > 0.98
    assert_batch_parity(res.sol_x, res.sol_y, res.cpg_info, ora, TOL, stable=stable)
    assert (res.cpg_info.status == ora['status']).all()          # statuses agree even on the unstable ones
    assert (res.cpg_info.obj_val[kind == 3] == 1e30).all() and (res.cpg_info.obj_val[kind == 4] == -1e30).all()


@pytest.mark.gpu
def test_shared_matrix_parameter_update():
    """A user parameter that enters a canonical MATRIX is shared by the batch: changing it re-runs the offline setup on
    the host (scaling, factor, schedule values) and re-uploads every constants table -- the role of
    osqp_update_data_mat in the reference's update tree (cvxpygen/solvers/osqp.py:20-33)."""
    import shutil, tempfile, os
    from cvxpygen_b200 import cpg, families, runtime
    from helpers import canon_batches
    d = os.path.join(tempfile.mkdtemp(), 'ls')
    fam = families.nonneg_ls(3, 2)
    cpg.generate_code(fam, code_dir=d, batch_params=['b'])
    mod = runtime.Module(d)
    B = 64
    rng = np.random.default_rng(3)
    bb = rng.standard_normal((B, 3))
    A_new = np.asarray(fam.param('A').default) * np.array([1.5, -0.7, 2.0]) + 0.1
    mod.update_shared_params({'A': A_new})
    res = mod.solve_batch({'b': bb}, return_canonical=True)
    fam2 = families.nonneg_ls(3, 2, A_data=A_new)
    q, l, u = canon_batches(fam2, {'b': bb}, B)
    ora = oracle_solve(fam2, q, l, u)
    assert_batch_parity(res.sol_x, res.sol_y, res.cpg_info, ora, TOL)
    with pytest.raises(ValueError):
        mod.update_shared_params({'b': np.zeros(3)})
    shutil.rmtree(os.path.dirname(d), ignore_errors=True)


@pytest.mark.gpu
def test_host_entry_zero_copy_equals_staged():
    """cpg_solve_batch_host: result rows stored by the kernels straight into PINNED host buffers (zero-copy) are bit-identical
    to the staged path (host_zero_copy = 0) and to pageable buffers; canonical x / y outputs included."""
    import torch
    name, B = 'mpc_12_4_10', 5000
    mod = standard.load(name)
    d = mod.dims
    xi = torch.from_numpy(np.random.default_rng(5).uniform(-1, 1, (B, 12)))

    def buffers(pinned):
        mk = lambda shape, dt=torch.float64: (torch.zeros(shape, dtype=dt).pin_memory() if pinned else torch.zeros(shape, dtype=dt))
        return dict(prim=mk((B, d.n_prim)), dual=mk((B, d.n_dual)), sol_x=mk((B, d.n_var)), sol_y=mk((B, d.n_con)),
                    obj=mk((B,)), pri=mk((B,)), dua=mk((B,)), it=mk((B,), torch.int32), st=mk((B,), torch.int32))
    hp = xi.clone().pin_memory()
    zc = mod.solve_batch_pinned(hp, buffers(True))
    mod.set_solver_setting('host_zero_copy', 0)
    staged = mod.solve_batch_pinned(hp, buffers(True))
    mod.set_solver_setting('host_zero_copy', 1)
    pageable = mod.solve_batch_pinned(xi, buffers(False))
    assert (zc['st'] == 1).all() and zc['prim'].abs().sum() > 0
    for k in zc:
        assert torch.equal(zc[k], staged[k]) and torch.equal(zc[k], pageable[k]), k
    with pytest.raises(AttributeError):
        mod.set_solver_setting('scaling', 3)


@pytest.mark.gpu
@pytest.mark.parametrize('adaptive_rho', [0, 1])
def test_big_family_parity(adaptive_rho):
    """f3: a family larger than one SM's shared memory (1 500-row KKT, tile schedule 530 KB).  The reference generates code for any
    size (cvxpygen/cpg.py:17-30); here such a family is solved by the per-instance-factor kernel alone (CPG_FAM_BIG) -- same
    iteration counts and statuses as the compiled reference on every instance."""
    import time
    name, B = 'random_qp_700_100_700', 192
    fam, params, (q, l, u) = family_and_batch(name, B, seed=7)
    mod = standard.load(name)
    res = mod.solve_batch(params, return_canonical=True, adaptive_rho=adaptive_rho)
    t0 = time.perf_counter()
    res = mod.solve_batch(params, return_canonical=True, adaptive_rho=adaptive_rho)
    dt = time.perf_counter() - t0
    mod.set_solver_default_settings()
    ora = oracle_solve(fam, q, l, u, adaptive_rho=adaptive_rho)
    assert_batch_parity(res.sol_x, res.sol_y, res.cpg_info, ora, TOL)
    assert mod.launch_count() == 2          # queue_all_kernel + admm_tail_kernel, no main kernel
    print(f'\nbig family {name}: B={B} host-call {dt * 1e3:.1f} ms = {B / dt:.0f} inst/s, mean iter {res.cpg_info.iter.mean():.1f}')


@pytest.mark.gpu
def test_tensor_core_main_kernel_parity():
    """The opt-in FP64 tensor-core main kernel (admm_dmma_kernel: mma.sync.m8n8k4.f64, eight instances per group of four warps;
    DESIGN 4.7) against the compiled reference: identical iteration counts and statuses, incl. rho updates handed to the tail kernel."""
    name, B = 'mpc_6_3_10_dmma', 700
    fam, params, (q, l, u) = family_and_batch(name, B, seed=9)
    mod = standard.load(name)
    hdr = open(__import__('os').path.join(standard.code_dir(name), 'c', 'include', 'cpg_family.h')).read()
    assert '#define CPG_FAM_DMMA 1' in hdr
    for kw in ({}, dict(adaptive_rho_interval=25, eps_abs=1e-5, eps_rel=1e-5)):
        res = mod.solve_batch(params, return_canonical=True, **kw)
        mod.set_solver_default_settings()
        ora = oracle_solve(fam, q, l, u, **kw)
        assert_batch_parity(res.sol_x, res.sol_y, res.cpg_info, ora, TOL, eps_abs=kw.get('eps_abs', 1e-3))
