"""Per-instance matrix parameters: the osqp_update_data_mat branch of the generated solve
(cvxpygen/solvers/osqp.py:20-33 -> osqp_update_P_A, osqp_sources/src/osqp.c:1158-1264) for a whole batch.

CPU part: the numpy restatement against the compiled reference; the kernel's tables (offline/blob.py:pack_matpar_blob)
emulated in numpy against a direct equilibration / KKT assembly.  GPU part: the CUDA path against the reference."""
import numpy as np
import pytest
import scipy.sparse as sp

from cvxpygen_b200 import families, standard
from cvxpygen_b200.offline import kkt as _kkt
from cvxpygen_b200.offline.equilibrate import ruiz_equilibrate
from cvxpygen_b200.offline.matpar import emulate_prepare, emulate_assemble
from cvxpygen_b200.offline.qp_setup import setup_qp_family, unscale_roundtrip
from cvxpygen_b200.offline.refactor import emulate_factor, emulate_solve

from helpers import ltv_batch, canon_matrix_batches, matrix_oracle_solve, ref_available, rel_err

BATCHED = ['A', 'B', 'qdiag', 'rdiag', 'x_init']


@pytest.fixture(scope='module')
def ltv_small():
    fam = families.mpc_ltv(4, 2, 5)
    return fam, setup_qp_family(fam, BATCHED)


def _theta_rows(fam, setup, params, B):
    th = np.tile(fam.theta_default(), (B, 1))
    for pn, v in params.items():
        p = fam.param(pn)
        th[:, p.col:p.col + p.size] = v
    return th[:, setup.batch_cols]


def test_pattern_is_structural():
    fam = families.mpc_ltv(4, 2, 5)
    assert fam.canon_matrix('A').nnz == (5 + 1) * 4 + 5 * 4 * (4 + 2) + 5 * 2          # dense A, B blocks stay in the pattern
    g = families.mpc(4, 2, 5)
    assert abs(fam.canon_matrix('A').toarray() - g.canon_matrix('A').toarray()).max() == 0.0
    assert abs(fam.canon_matrix('P').toarray() - g.canon_matrix('P').toarray()).max() == 0.0


def test_tables_equilibrate_bit_exact(ltv_small):
    """matpar_prepare's operation sequence on the packed tables == scale_data on the instance's matrices, bit for bit."""
    fam, setup = ltv_small
    B = 4
    params = ltv_batch(fam, B)
    Px, Ax, _ = canon_matrix_batches(fam, params, B)
    thb = _theta_rows(fam, setup, params, B)
    sc0 = ruiz_equilibrate(fam.canon_matrix('P'), fam.canon_matrix('A'), fam.canon_data('q'), 10)
    q_un = unscale_roundtrip(sc0, 10)[2]
    Pi, Pp, Ps = fam.patterns['P']; Ai, Ap, As = fam.patterns['A']
    for i in range(B):
        emu = emulate_prepare(setup.mat_blob, thb[i])
        ref = ruiz_equilibrate(sp.csc_matrix((Px[i], Pi, Pp), shape=Ps), sp.csc_matrix((Ax[i], Ai, Ap), shape=As), q_un, 10)
        assert np.array_equal(emu['D'], ref['D']) and np.array_equal(emu['E'], ref['E']) and emu['c'] == ref['c']
        assert np.array_equal(emu['Pv'], ref['P'].data) and np.array_equal(emu['Av'], ref['A'].data)


def test_tables_assemble_and_factor(ltv_small):
    """slot maps: the assembled S is the permuted KKT matrix; factor + solve on it == dense solve."""
    fam, setup = ltv_small
    params = ltv_batch(fam, 2, seed=5)
    thb = _theta_rows(fam, setup, params, 2)
    n, m = setup.n, setup.m
    rho_vec = _kkt.rho_vector(setup.ctype, setup.rho)
    Pi, Pp, Ps = fam.patterns['P']; Ai, Ap, As = fam.patterns['A']
    for i in range(2):
        emu = emulate_prepare(setup.mat_blob, thb[i])
        S = emulate_assemble(emu['blob'], emu['Pv'], emu['Av'], setup.refactor.rho_slot, rho_vec)
        K = _kkt.assemble_kkt(sp.csc_matrix((emu['Pv'], Pi, Pp), shape=Ps), sp.csc_matrix((emu['Av'], Ai, Ap), shape=As),
                              setup.sigma, rho_vec).toarray()
        perm = setup.factor.perm
        Kp = K[np.ix_(perm, perm)]
        T = setup.refactor
        assert np.array_equal(S[:n + m], np.diag(Kp))
        ii, jj = np.nonzero(T.slot >= 0)
        assert np.array_equal(S[T.slot[ii, jj]], Kp[ii, jj])
        # numeric factorisation + solve with the kernel's op lists on these values
        T2 = type(T)(**{**T.__dict__, 'S0': np.where(np.isin(np.arange(T.n_slots), T.rho_slot), 0.0, S)})
        Sf = emulate_factor(T2, rho_vec)
        rhs = np.random.default_rng(i).standard_normal(n + m)
        got = emulate_solve(T2, Sf, rhs[perm])
        want = np.linalg.solve(K, rhs)[perm]
        assert np.abs(got - want).max() < 1e-9 * max(1.0, np.abs(want).max())


@pytest.mark.skipif(not ref_available(), reason='oracle/_ref not built')
def test_numpy_restatement_vs_reference_matrix_updates():
    """oracle pinning for this path: osqp_update_P_A + update_lin_cost/bounds + osqp_solve (compiled 0.6.2) vs the numpy
    restatement -- identical iteration counts and statuses, x / y to 1e-9."""
    fam = families.mpc_ltv(6, 3, 10)
    B = 6
    params = ltv_batch(fam, B)
    Px, Ax, (q, l, u) = canon_matrix_batches(fam, params, B)
    ref = matrix_oracle_solve(fam, Px, Ax, q, l, u, prefer_ref=True)
    npy = matrix_oracle_solve(fam, Px, Ax, q, l, u, prefer_ref=False)
    assert np.array_equal(ref['iter'], npy['iter']) and np.array_equal(ref['status'], npy['status'])
    assert (ref['status'] == 1).all()
    assert rel_err(npy['x'], ref['x']).max() < 1e-9 and rel_err(npy['y'], ref['y']).max() < 1e-9
    # and the matrices matter: the shared-matrix solve of the same x_init gives another answer
    from helpers import oracle_solve
    shared = oracle_solve(fam, q, l, u)
    assert rel_err(shared['x'], ref['x']).min() > 1e-3


def test_generated_library_exports_matrix_entry():
    import ctypes, os
    d = standard.build('mpc_ltv_6_3_10')
    lib = ctypes.CDLL(os.path.join(d, 'libcpg_b200.so'))
    for sym in ('cpg_b200_load_mat_constants', 'cpg_solve_batch_device', 'cpg_solve_batch_host'):
        assert hasattr(lib, sym)
    hdr = open(os.path.join(d, 'c', 'include', 'cpg_family.h')).read()
    assert '#define CPG_FAM_MATPAR 1' in hdr


def test_pack_params_matrix_layouts():
    """Host side of the batched call for matrix parameters: per-instance matrices are flattened in Fortran order (reference
    TPL/cpg_solver.py.jinja2:26-34), one value broadcasts over the batch, sparse / diagonal parameters are passed as their
    stored entries, missing parameters take their defaults."""
    from cvxpygen_b200 import runtime
    mod = runtime.load(standard.build('mpc_ltv_6_3_10'), 0)            # no device needed for packing
    fam = standard.STANDARD['mpc_ltv_6_3_10'][0]()
    B = 5
    rng = np.random.default_rng(0)
    A = rng.standard_normal((B, 6, 6)); Bm = rng.standard_normal((6, 3)); xi = rng.standard_normal((B, 6))
    rows = mod.pack_params({'A': A, 'B': Bm, 'x_init': xi})
    assert rows.shape == (B, mod.dims.n_param) == (B, 36 + 18 + 6 + 3 + 6)
    cols = {p.name: (p.col, p.size) for p in fam.params}
    off = lambda nm: sum(fam.param(x).size for x in standard.STANDARD['mpc_ltv_6_3_10'][1][:standard.STANDARD['mpc_ltv_6_3_10'][1].index(nm)])
    for b in range(B):
        assert np.array_equal(rows[b, off('A'):off('A') + 36], A[b].flatten(order='F'))
        assert np.array_equal(rows[b, off('B'):off('B') + 18], Bm.flatten(order='F'))            # broadcast
        assert np.array_equal(rows[b, off('qdiag'):off('qdiag') + 6], fam.param('qdiag').default)   # default
        assert np.array_equal(rows[b, off('x_init'):off('x_init') + 6], xi[b])
    with pytest.raises(AttributeError):
        mod.pack_params({'nope': xi})
    with pytest.raises(ValueError):
        mod.pack_params({'A': A, 'x_init': xi[:3]})
    ref = runtime.load(standard.build('mpc_ref_6_3_10'), 0)
    r2 = ref.pack_params({'Qsqrt': np.full((B, 6), 2.0), 'A': np.arange(9.0)})      # stored entries of the diagonal / sparse parameters
    assert r2.shape == (B, 6 + 6 + 3 + 9 + 3 + 6) and np.array_equal(r2[:, 6:12], np.full((B, 6), 2.0))
    assert np.array_equal(r2[0, 15:24], np.arange(9.0))


def _kat_batches():
    fam, cases = families.osqp_update_matrices_kat()
    B = 4
    vec = {k: np.tile(fam.param(k).default, (B, 1)) for k in ('q', 'l', 'u')}
    return fam, cases, dict(vec, P=cases['P'], A=cases['A'])


def test_osqp_update_matrices_known_answers():
    """Oracle pinning for the matrix-update path on OSQP's OWN known-answer test (osqp_sources/tests/update_matrices: the
    data generator is repeated stream for stream, the expected values are the ones test_update_matrices.h asserts with
    TESTS_TOL = 1e-4): original problem, P updated, A updated, both -- compiled reference (when built) and numpy restatement."""
    fam, cases, params = _kat_batches()
    for prefer_ref in ([True, False] if ref_available() else [False]):
        o = matrix_oracle_solve(fam, cases['P'], cases['A'], params['q'], params['l'], params['u'], prefer_ref=prefer_ref, max_iter=1000)
        assert (o['status'] == 1).all()
        assert np.abs(o['x'] - cases['x']).max() < 1e-4 and np.abs(o['y']).max() < 1e-4
        assert np.abs(o['obj'] - cases['obj']).max() < 1e-4


def test_osqp_update_matrices_known_answers_on_the_kernel(tmp_path):
    """... and the matrix-parameter kernel itself (product source on the SIMT emulator) on the same four variants."""
    from test_simt_emulation import build_emu, run_solve, _rows
    fam, cases, params = _kat_batches()
    st, lib, dims = build_emu(fam, ['q', 'l', 'u', 'P', 'A'], str(tmp_path))
    out = run_solve(lib, 'emu_matpar_solve', dims, _rows(fam, st, params, 4), grid=1)
    assert (out['status'] == 1).all()
    assert np.abs(out['x'] - cases['x']).max() < 1e-4 and np.abs(out['y']).max() < 1e-4 and np.abs(out['obj'] - cases['obj']).max() < 1e-4


@pytest.mark.gpu
def test_gpu_osqp_update_matrices_known_answers():
    fam, cases, params = _kat_batches()
    res = standard.load('osqp_update_matrices_5_8', device=0).solve_batch(params, return_canonical=True)
    assert (res.cpg_info.status == 1).all()
    assert np.abs(res.sol_x - cases['x']).max() < 1e-4 and np.abs(res.sol_y).max() < 1e-4
    assert np.abs(res.cpg_info.obj_val - cases['obj']).max() < 1e-4
    ora = matrix_oracle_solve(fam, cases['P'], cases['A'], params['q'], params['l'], params['u'])
    assert np.array_equal(res.cpg_info.iter, ora['iter']) and np.abs(res.sol_x - ora['x']).max() < 1e-9


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize('name,B', [('mpc_ltv_6_3_10', 300), ('mpc_ltv_12_4_10', 200)])
def test_gpu_matrix_parameters_vs_reference(name, B):
    from cvxpygen_b200 import runtime
    fam = standard.STANDARD[name][0]()
    mod = standard.load(name, device=0)
    params = ltv_batch(fam, B, seed=11)
    res = mod.solve_batch(params, return_canonical=True)
    assert mod.launch_count() == 1
    Px, Ax, (q, l, u) = canon_matrix_batches(fam, params, B)
    ora = matrix_oracle_solve(fam, Px, Ax, q, l, u)
    info = res.cpg_info
    assert np.array_equal(info.status, ora['status']), np.nonzero(info.status != ora['status'])[0][:5]
    assert np.array_equal(info.iter, ora['iter']), np.nonzero(info.iter != ora['iter'])[0][:5]
    assert (info.status == 1).mean() > 0.95
    ok = np.isin(info.status, [1, 2, -2])
    assert rel_err(res.sol_x[ok], ora['x'][ok]).max() < 1e-7       # north_star tolerance is 1e-5
    assert rel_err(res.sol_y[ok], ora['y'][ok]).max() < 1e-7
    assert np.allclose(info.obj_val[ok], ora['obj'][ok], rtol=1e-7, atol=1e-10)
    # user-level gathers
    X = res.cpg_prim['X']
    assert np.abs(X[:, :, 0] - params['x_init']).max() < 1e-2      # x_0 = x_init to solver accuracy


@pytest.mark.gpu
def test_gpu_sparse_matrix_parameter_only_A_batched():
    """README example (examples/main.py:16-26) with its sparse parameter A per instance: only A is outdated, so the
    reference calls osqp_update_data_mat with P = NULL (cvxpygen/solvers/osqp.py:28-31) and P keeps its unscale/re-scale
    round-trip values -- the kernel's base table carries exactly those."""
    name, B = 'nonneg_LS_3_2_A', 400
    fam = standard.STANDARD[name][0]()
    rng = np.random.default_rng(8)
    params = {'A': fam.param('A').default[None, :] + 0.5 * rng.standard_normal((B, 3)),
              'b': fam.param('b').default[None, :] + 0.5 * rng.standard_normal((B, 3))}
    res = standard.load(name, device=0).solve_batch(params, return_canonical=True)
    _, Ax, (q, l, u) = canon_matrix_batches(fam, params, B)
    ora = matrix_oracle_solve(fam, None, Ax, q, l, u)
    assert np.array_equal(res.cpg_info.status, ora['status']) and np.array_equal(res.cpg_info.iter, ora['iter'])
    ok = np.isin(ora['status'], [1, 2, -2])
    assert ok.mean() > 0.9
    assert np.abs(res.sol_x[ok] - ora['x'][ok]).max() < 1e-7 * max(1.0, np.abs(ora['x'][ok]).max())
    assert np.abs(res.sol_y[ok] - ora['y'][ok]).max() < 1e-7 * max(1.0, np.abs(ora['y'][ok]).max())


def _reference_mpc_case(B, seed):
    fam = standard.STANDARD['mpc_ref_6_3_10'][0]()
    params = families.mpc_reference_batch(fam, B, seed=seed)
    _, Ax, (q, l, u) = canon_matrix_batches(fam, params, B)
    return fam, params, Ax, q, l, u


@pytest.mark.skipif(not ref_available(), reason='oracle/_ref not built')
def test_reference_test_mpc_all_parameters_oracles_agree():
    """The reference's own test MPC (tests/test_E2E_QP.py:44-73) with its six parameters per instance: diagonal cost
    factors, sparse A and B (stored entries in column-major order), x_init.  P is constant -> only-A branch."""
    fam, params, Ax, q, l, u = _reference_mpc_case(5, seed=0)
    assert fam.n_var == 192 and fam.param('A').size == 9 and fam.param('B').size == 3 and fam.param('Qsqrt').size == 6
    ref = matrix_oracle_solve(fam, None, Ax, q, l, u)
    npy = matrix_oracle_solve(fam, None, Ax, q, l, u, prefer_ref=False)
    assert np.array_equal(ref['iter'], npy['iter']) and (ref['status'] == 1).all()
    assert rel_err(npy['x'], ref['x']).max() < 1e-8
    # the canonical form really is the reference's problem: dynamics hold and the objective is the user-level one (minus d = 1)
    tight = matrix_oracle_solve(fam, None, Ax, q, l, u, eps_abs=1e-9, eps_rel=1e-9)
    X = tight['x'][:, :66].reshape(5, 11, 6); U = tight['x'][:, 66:96].reshape(5, 10, 3)
    obj = ((params['Psqrt'] * X[:, 9]) ** 2).sum(1) + ((params['Qsqrt'][:, None, :] * X[:, :10]) ** 2).sum((1, 2)) \
        + ((params['Rsqrt'][:, None, :] * U) ** 2).sum((1, 2))
    assert np.allclose(obj, tight['obj'], rtol=1e-6)
    assert np.abs(X[:, 0] - params['x_init']).max() < 1e-7 and np.abs(U).max() <= 1 + 1e-7


@pytest.mark.gpu
def test_gpu_reference_test_mpc_all_parameters():
    B = 1000
    fam, params, Ax, q, l, u = _reference_mpc_case(B, seed=1)
    res = standard.load('mpc_ref_6_3_10', device=0).solve_batch(params, return_canonical=True)
    ora = matrix_oracle_solve(fam, None, Ax, q, l, u)
    assert np.array_equal(res.cpg_info.status, ora['status']) and np.array_equal(res.cpg_info.iter, ora['iter'])
    assert (ora['status'] == 1).mean() > 0.99
    ok = ora['status'] == 1
    assert rel_err(res.sol_x[ok], ora['x'][ok]).max() < 1e-7 and rel_err(res.sol_y[ok], ora['y'][ok]).max() < 1e-7
    assert np.allclose(res.cpg_info.obj_val[ok], ora['obj'][ok] + 1.0, rtol=1e-7)      # cpg_retrieve_info adds d = 1 (utils.py:980)
    assert np.abs(res.cpg_prim['X'][:, :, 0] - params['x_init']).max() < 2e-2


@pytest.mark.gpu
def test_gpu_matrix_parameters_warm_start():
    """x0 / y0 (osqp_warm_start, osqp.c:929-953) on the matrix-parameter path: starting from the solution stops at the first check."""
    B = 300
    fam, params, Ax, q, l, u = _reference_mpc_case(B, seed=2)
    mod = standard.load('mpc_ref_6_3_10', device=0)
    cold = mod.solve_batch(params, return_canonical=True)
    warm = mod.solve_batch(params, x0=cold.sol_x, y0=cold.sol_y, return_canonical=True)
    assert (warm.cpg_info.iter == 25).all() and (warm.cpg_info.status == 1).all()
    assert rel_err(warm.sol_x, cold.sol_x).max() < 1e-2


@pytest.mark.gpu
def test_gpu_shared_parameter_update_on_matrix_family():
    """A family generated with only A, x_init batched: B and the cost factors are SHARED; changing them = host re-setup +
    cpg_b200_load_constants_all + cpg_b200_load_mat_constants, then the batch is solved with the new shared values."""
    import tempfile
    from cvxpygen_b200 import cpg, runtime
    fam = families.mpc_reference(10)
    with tempfile.TemporaryDirectory() as td:
        cpg.generate_code(fam, code_dir=td, solver='ADMM-CUDA', batch_params=['A', 'x_init'])
        mod = runtime.load(td, 0)
        B = 64
        full = families.mpc_reference_batch(fam, B, seed=5)
        newB = fam.param('B').default * 1.3
        newR = fam.param('Rsqrt').default * 0.7
        mod.update_shared_params({'B': newB, 'Rsqrt': newR})
        res = mod.solve_batch({'A': full['A'], 'x_init': full['x_init']}, return_canonical=True)
        # reference: the same family object with the new defaults, all instances sharing B / Rsqrt
        fam2 = families.mpc_reference(10)
        fam2.param('B').default[:] = newB; fam2.param('Rsqrt').default[:] = newR
        p2 = {'A': full['A'], 'x_init': full['x_init']}
        _, Ax, (q, l, u) = canon_matrix_batches(fam2, p2, B)
        ora = matrix_oracle_solve(fam2, None, Ax, q, l, u)
        assert np.array_equal(res.cpg_info.iter, ora['iter']) and np.array_equal(res.cpg_info.status, ora['status'])
        assert rel_err(res.sol_x, ora['x']).max() < 1e-7


@pytest.mark.skipif(not ref_available(), reason='oracle/_ref not built')
def test_reference_actuator_problem_oracles_agree():
    """The reference's "degenerate vectors and matrices" test problem (tests/test_E2E_QP.py:16-41): one actuator, a 1 x 1
    matrix variable, scalar parameters; lamb_sm enters P and A enters the constraint matrix (osqp_update_P_A branch)."""
    fam = families.actuator()
    B = 32
    params = families.actuator_batch(fam, B, seed=1)
    Px, Ax, (q, l, u) = canon_matrix_batches(fam, params, B)
    ref = matrix_oracle_solve(fam, Px, Ax, q, l, u)
    npy = matrix_oracle_solve(fam, Px, Ax, q, l, u, prefer_ref=False)
    assert np.array_equal(ref['iter'], npy['iter']) and (ref['status'] == 1).all()
    assert np.abs(ref['x'] - npy['x']).max() < 1e-7
    tight = matrix_oracle_solve(fam, Px, Ax, q, l, u, eps_abs=1e-10, eps_rel=1e-10, max_iter=100000)
    uu = tight['x'][:, 0]
    obj = ((params['A'] * uu[:, None] - params['w']) ** 2).sum(1) + params['lamb_sm'][:, 0] * (uu - params['u_prev'][:, 0]) ** 2 \
        + params['kappa'][:, 0] * np.abs(uu)
    assert np.allclose(obj, tight['obj'], rtol=1e-8, atol=1e-9)
    st = setup_qp_family(fam, standard.STANDARD['actuator_1_3'][1])
    assert st.mat_params == ['A', 'lamb_sm']


@pytest.mark.gpu
def test_gpu_reference_actuator_problem():
    name, B = 'actuator_1_3', 2000
    fam = standard.STANDARD[name][0]()
    params = families.actuator_batch(fam, B, seed=2)
    mod = standard.load(name, device=0)
    res = mod.solve_batch(params, return_canonical=True)
    Px, Ax, (q, l, u) = canon_matrix_batches(fam, params, B)
    ora = matrix_oracle_solve(fam, Px, Ax, q, l, u)
    assert np.array_equal(res.cpg_info.status, ora['status']) and np.array_equal(res.cpg_info.iter, ora['iter'])
    ok = ora['status'] == 1
    assert ok.mean() > 0.99 and np.abs(res.sol_x[ok] - ora['x'][ok]).max() < 1e-7 and np.abs(res.sol_y[ok] - ora['y'][ok]).max() < 1e-6
    assert res.cpg_prim['u'].shape == (B, 1) and res.cpg_prim['delta_u'].shape == (B, 1, 1)
    assert np.abs(res.cpg_prim['delta_u'][:, 0, 0] - (res.cpg_prim['u'][:, 0] - params['u_prev'][:, 0]))[ok].max() < 1e-2
    # backward pass with P AND A per instance: d/dtheta of u against the numpy restatement
    dprim = np.zeros((B, 2)); dprim[:, 0] = 1.0
    g = mod.gradient_batch_mat(params, res.sol_x, res.sol_y, dprim)
    nchk = 64
    dx = np.zeros((nchk, fam.n_var)); dx[:, 0] = 1.0
    dq, dl, du, dP, dA = qp_backward_mat(fam.patterns['P'], fam.patterns['A'], Px[:nchk], Ax[:nchk], res.sol_x[:nchk], res.sol_y[:nchk], dx)
    names = standard.STANDARD[name][1]
    want = param_gradient_mat(fam, dq, dl, du, dP, dA, names)
    got = np.concatenate([g[nm][:nchk] for nm in names], axis=1)
    # u = 0 makes BOTH rows of |u| <= t active: K is then singular up to the 1e-6 regularisation and the answer depends on the
    # factorisation at the 1e-4 level (LDL' on the symbolic pattern here, pivoted LU in the restatement, up/down-dated LDL' in
    # the reference) -- those instances get the looser bar
    y = res.sol_y[:nchk]
    degenerate = (np.abs(y[:, 5]) > 1e-12) & (np.abs(y[:, 6]) > 1e-12)
    assert (~degenerate).mean() > 0.7
    scale = max(1.0, np.abs(want).max())
    assert np.abs(got - want)[~degenerate].max() < 1e-6 * scale
    assert np.abs(got - want).max() < 1e-3 * scale


@pytest.mark.gpu
def test_gpu_matrix_parameters_reduce_to_shared_family():
    """With every instance carrying the DEFAULT matrices the matrix-parameter kernel must reproduce the shared-matrix
    kernels' results on the same x_init (same algorithm, per-instance factor instead of the family's)."""
    B = 500
    fam = families.mpc_ltv(12, 4, 10)
    xi = np.random.default_rng(2).uniform(-1, 1, (B, 12))
    a = standard.load('mpc_ltv_12_4_10', device=0).solve_batch({'x_init': xi}, return_canonical=True)
    b = standard.load('mpc_12_4_10', device=0).solve_batch({'x_init': xi}, return_canonical=True)
    assert np.array_equal(a.cpg_info.iter, b.cpg_info.iter) and np.array_equal(a.cpg_info.status, b.cpg_info.status)
    assert rel_err(a.sol_x, b.sol_x).max() < 1e-7 and rel_err(a.sol_y, b.sol_y).max() < 1e-7


@pytest.mark.gpu
def test_gpu_matrix_parameters_large_batch_properties():
    """20 000 instances: optimality conditions of every instance's OWN problem (unscaled), run-to-run reproducibility."""
    name, B = 'mpc_ltv_12_4_10', 20000
    fam = standard.STANDARD[name][0]()
    mod = standard.load(name, device=0)
    params = ltv_batch(fam, B, seed=21)
    r1 = mod.solve_batch(params, return_canonical=True)
    r2 = mod.solve_batch(params, return_canonical=True)
    assert np.array_equal(r1.sol_x, r2.sol_x, equal_nan=True) and np.array_equal(r1.cpg_info.iter, r2.cpg_info.iter)
    ok = r1.cpg_info.status == 1
    assert ok.mean() > 0.99
    Px, Ax, (q, l, u) = canon_matrix_batches(fam, params, B)
    Ai, Ap, As = fam.patterns['A']
    rows = np.asarray(Ai); cols = np.repeat(np.arange(As[1]), np.diff(Ap))
    Axx = np.zeros((B, As[0]))
    np.add.at(Axx, (slice(None), rows), Ax * r1.sol_x[:, cols])
    viol = np.maximum(np.maximum(l - Axx, Axx - u), 0.0).max(axis=1)
    scale = np.maximum(np.abs(Axx).max(axis=1), 1.0)
    assert (viol[ok] < 2e-3 * (1 + scale[ok])).all()               # eps_abs + eps_rel * |Ax| with eps = 1e-3
    ATy = np.zeros((B, As[1]))
    np.add.at(ATy, (slice(None), cols), Ax * r1.sol_y[:, rows])
    Pdiag = Px                                                      # P is diagonal in this family
    rd = np.abs(Pdiag * r1.sol_x + q + ATy).max(axis=1)
    sc2 = np.maximum(np.maximum(np.abs(Pdiag * r1.sol_x).max(axis=1), np.abs(ATy).max(axis=1)), 1.0)
    assert (rd[ok] < 2e-3 * (1 + sc2[ok])).all()


# ------------------------------------------------------------------------------------------------ backward pass (a16 + f2)
import os
from helpers import GOLDEN
from oracle.grad_numpy import qp_backward_mat, param_gradient_mat


def _relmax(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def test_numpy_matrix_backward_matches_reference_generated_c():
    """numpy restatement vs the reference's generated C on per-instance matrices (cpg_P_to_K / cpg_A_to_K +
    cpg_ldl_numeric + cpg_osqp_gradient), golden vectors from tests/golden/make_golden_grad.py."""
    name = 'mpc_ltv_6_3_10'
    g = np.load(os.path.join(GOLDEN, f'grad_mat_{name}.npz'))
    fam = standard.STANDARD[name][0]()
    prim_idx = np.concatenate([v.indices for v in fam.variables])
    dx = np.zeros((g['dprim'].shape[0], fam.n_var)); dx[:, prim_idx] = g['dprim']
    dq, dl, du, dP, dA = qp_backward_mat(fam.patterns['P'], fam.patterns['A'], g['Px'], g['Ax'], g['sol_x'], g['sol_y'], dx)
    assert _relmax(dq, g['dq']) < 1e-9 and _relmax(dl + du, g['dl'] + g['du']) < 1e-9
    assert _relmax(dP, g['dP']) < 1e-9 and _relmax(dA, g['dA']) < 1e-9


@pytest.mark.skipif(not ref_available(), reason='oracle/_ref not built')
def test_matrix_backward_matches_finite_differences():
    """d/dtheta of c'x*(theta) w.r.t. entries of A, B, qdiag, rdiag, x_init of one instance (tight forward solves)."""
    fam = families.mpc_ltv(4, 2, 5)
    B = 2
    params = ltv_batch(fam, B, seed=13)
    kw = dict(eps_abs=1e-11, eps_rel=1e-11, max_iter=400000)

    def solve(pr):
        Px, Ax, (q, l, u) = canon_matrix_batches(fam, pr, B)
        return matrix_oracle_solve(fam, Px, Ax, q, l, u, **kw), Px, Ax
    sol, Px, Ax = solve(params)
    cvec = np.random.default_rng(1).standard_normal(fam.n_var)
    dx = np.tile(cvec, (B, 1))
    dq, dl, du, dP, dA = qp_backward_mat(fam.patterns['P'], fam.patterns['A'], Px, Ax, sol['x'], sol['y'], dx)
    names = ['A', 'B', 'qdiag', 'rdiag', 'x_init']
    g = param_gradient_mat(fam, dq, dl, du, dP, dA, names)
    cols = fam.param_columns(names)
    h = 1e-6
    rng = np.random.default_rng(4)
    for k in rng.choice(len(cols), 12, replace=False):
        nm = next(p.name for p in fam.params if p.col <= cols[k] < p.col + p.size)
        off = cols[k] - fam.param(nm).col
        fd = np.zeros(B)
        for sgn in (+1, -1):
            p2 = {a: v.copy() for a, v in params.items()}
            p2[nm][:, off] += sgn * h
            fd += sgn * (solve(p2)[0]['x'] @ cvec) / (2 * h)
        assert np.allclose(g[:, k], fd, rtol=2e-3, atol=2e-5), (nm, off, g[:, k], fd)


@pytest.mark.gpu
def test_gpu_matrix_backward_matches_reference_golden():
    name = 'mpc_ltv_6_3_10'
    g = np.load(os.path.join(GOLDEN, f'grad_mat_{name}.npz'))
    fam = standard.STANDARD[name][0]()
    mod = standard.load(name)
    names = standard.STANDARD[name][1]
    params = {nm: g['param_' + nm] for nm in names}
    res, dq, dl, du, dP, dA = mod.gradient_batch_mat(params, g['sol_x'], g['sol_y'], g['dprim'], return_canonical=True)
    assert mod.launch_count() == 1
    assert _relmax(dq, g['dq']) < 1e-5 and _relmax(dl + du, g['dl'] + g['du']) < 1e-5
    assert _relmax(dP, g['dP']) < 1e-5 and _relmax(dA, g['dA']) < 1e-5
    ref = param_gradient_mat(fam, g['dq'], g['dl'], g['du'], g['dP'], g['dA'], names)
    got = np.concatenate([res[nm] for nm in names], axis=1)
    assert _relmax(got, ref) < 1e-5
    with pytest.raises(RuntimeError):          # the shared-matrix entry has no parameter rows to canonicalise P / A from
        mod.gradient_batch(g['sol_y'], g['dprim'])


@pytest.mark.gpu
def test_gpu_matrix_forward_backward_torch_layer():
    import torch
    from cvxpygen_b200.torch_layer import BatchedQPLayer
    name, B = 'mpc_ltv_12_4_10', 256
    fam = standard.STANDARD[name][0]()
    mod = standard.load(name)
    params = ltv_batch(fam, B, seed=41)
    th = torch.tensor(mod.pack_params(params), dtype=torch.float64, device='cuda', requires_grad=True)
    prim = BatchedQPLayer(mod)(th)
    wgt = torch.randn(prim.shape, dtype=torch.float64, device='cuda', generator=torch.Generator('cuda').manual_seed(0))
    (prim * wgt).sum().backward()
    nchk = 32
    sub = {k: v[:nchk] for k, v in params.items()}
    Px, Ax, (q, l, u) = canon_matrix_batches(fam, sub, nchk)
    sol = matrix_oracle_solve(fam, Px, Ax, q, l, u)
    prim_idx = np.concatenate([v.indices for v in fam.variables])
    assert np.allclose(prim[:nchk].detach().cpu().numpy(), sol['x'][:, prim_idx], rtol=1e-6, atol=1e-9)
    dx = np.zeros((nchk, fam.n_var)); dx[:, prim_idx] = wgt[:nchk].cpu().numpy()
    dq, dl, du, dP, dA = qp_backward_mat(fam.patterns['P'], fam.patterns['A'], Px, Ax, sol['x'], sol['y'], dx)
    ref = param_gradient_mat(fam, dq, dl, du, dP, dA, standard.STANDARD[name][1])
    assert _relmax(th.grad[:nchk].cpu().numpy(), ref) < 1e-5
