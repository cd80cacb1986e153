"""IPM-CUDA backend (BASELINE config 3: portfolio SOCP through the Mehrotra interior-point kernel).

CPU tests: the offline tables (equilibration, table-driven LDL' and triangular solves), the phase logic of the kernel
compiled as host C++ (tests/emu, every barrier-separated phase run thread after thread) against the golden vectors of
the compiled reference, the generated directory and the exported C ABI.
GPU tests (-m gpu): the CUDA kernel through the C ABI against the golden vectors and against the compiled reference
on fresh seeded batches, bit-level properties at full batch size."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import scipy.sparse as sp

from cvxpygen_b200 import families, standard, codegen_ipm, runtime
from cvxpygen_b200.offline import socp_setup as ss
from helpers import GOLDEN
from oracle import ref_ecos

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAME = 'portfolio_socp_100_10'
# parity bar of the north star: 1e-5 relative on primal / dual variables.  What is actually reached against the compiled
# reference is ~1e-9 on x, y, s and ~1e-7 on z (dual degenerate epigraph rows amplify rounding differences).
RTOL_PRIMAL, RTOL_DUAL = 1e-7, 1e-5


@pytest.fixture(scope='module')
def setup():
    return ss.setup_socp_family(families.portfolio_socp())


def _golden():
    return np.load(os.path.join(GOLDEN, 'socp_portfolio_100_10.npz'))


def _rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


# ---------------------------------------------------------------------------------------------------------------------
def test_equilibration_matches_reference_restatement(setup):
    """cvxpygen_b200.offline.socp_setup.ecos_equilibrate against the oracle's restatement of ecos/src/equil.c."""
    from oracle.ipm_numpy import ruiz_equilibrate
    fam = setup.family
    A, G, xe, Ae, Ge = ruiz_equilibrate(fam.canon_matrix('A'), fam.canon_matrix('G'), setup.l, setup.q)
    assert np.array_equal(xe, setup.xe) and np.array_equal(Ae, setup.Ae) and np.array_equal(Ge, setup.Ge)
    assert np.array_equal(A.toarray(), setup.A_eq.toarray()) and np.array_equal(G.toarray(), setup.G_eq.toarray())


def test_structure_of_the_schedule(setup):
    D = setup.defines
    assert (D['N'], D['P'], D['M'], D['L'], D['NSOC'], D['NK']) == (512, 111, 715, 601, 2, 1342)
    assert D['NT'] <= ss.MAX_TAIL and D['NLW'] >= 1
    assert np.all(np.diff(setup.pos_level) >= 0)
    # every wide level's columns are mutually independent: no L entry inside a level
    T = setup.tables
    inv = np.empty(D['NK'], int); inv[T['perm']] = np.arange(D['NK'])
    lev_of_k = setup.pos_level[inv]
    assert np.all(lev_of_k[T['bw_s']] > lev_of_k[T['bw_t']])


def test_gather_plan_sums_every_row_once():
    """cvxpygen_b200.offline.gather: ragged rows (empty, single, longer than a warp's quota) dealt over 64 threads; every
    target is committed exactly once with the sum of its entries, flags and group alignment survive the packing."""
    from cvxpygen_b200.offline import gather
    rs = np.random.RandomState(3)
    vals = rs.randn(500)
    sizes = [0, 1, 2, 3, 7, 40, 150, 0, 5, 1] * 9
    plan = gather.GatherPlan(T=64, nfields=2, tbits=11, null_entry=(len(vals), 0))
    rows = []
    for t, c in enumerate(sizes):
        rows.append((t, t % 2, [(int(i), int(rs.randint(0, 100))) for i in rs.randint(0, len(vals), c)]))
    gather.add_phase(plan, rows[:30]); gather.add_phase(plan, []); gather.add_phase(plan, rows[30:])
    assert [p.round_hi - p.round_lo for p in plan.phases][1] == 0
    v = np.r_[vals, 0.0]
    got = {}

    def commit(t, flag, acc):
        assert t not in got and flag == t % 2
        got[t] = acc
    for ph in range(3):
        gather.run_phase(plan, ph, lambda e: v[e[0]], commit)
    assert sorted(got) == list(range(len(sizes)))
    for t, _, ent in rows:
        assert abs(got[t] - sum(vals[i] for i, _ in ent)) < 1e-12
    assert plan.desc_array().dtype == np.uint16 and plan.entry_array().shape[1] == 2
    assert len(plan.wr_base) == plan.n_rounds * 2 and max(plan.wr_shuf) >= 1      # the long rows are split over lanes


def test_table_driven_factor_and_solve_match_dense_algebra(setup):
    D, T = setup.defines, setup.tables
    rs = np.random.RandomState(0)
    nk, zoff, mt = D['NK'], D['ZOFF'], D['MT']
    _, blocks, _ = ss.stretch_layout(setup.l, setup.q)
    zd = -(0.5 + rs.rand(mt))
    sign = np.r_[np.ones(setup.n), -np.ones(setup.p), -np.ones(mt)]
    for o, so, d in blocks:
        zd[so + d + 1] = 0.7; sign[zoff + so + d + 1] = 1
    sv, su = 0.1 * rs.randn(len(T['socv'])), 0.1 * rs.randn(len(T['socu']))
    K = np.zeros((nk, nk))
    K[np.arange(setup.n), np.arange(setup.n)] = ss.DELTASTAT
    K[setup.n + np.arange(setup.p), setup.n + np.arange(setup.p)] = -ss.DELTASTAT
    for r, c, v in zip(T['mr_t'], T['mr_s'], T['ag_val']):
        K[r, c] += v; K[c, r] += v
    K[zoff + np.arange(mt), zoff + np.arange(mt)] = zd
    e = f = 0
    for o, so, d in blocks:
        iv, iu = zoff + so + d, zoff + so + d + 1
        for r in range(1, d):
            K[zoff + so + r, iv] = K[iv, zoff + so + r] = sv[e]; e += 1
        for r in range(d):
            K[zoff + so + r, iu] = K[iu, zoff + so + r] = su[f]; f += 1
    S, Dinv = ss.emulate_factor(setup, ss.fill_slots(setup, zd, sv, su), sign)
    b = rs.randn(nk)
    x = ss.emulate_solve(setup, S, Dinv, b)
    assert np.abs(K @ x - b).max() < 1e-6 * np.abs(b).max()         # static regularisation 7e-8 limits the accuracy
    assert _rel(x, np.linalg.solve(K, b)) < 1e-7


def test_exact_restatement_reproduces_the_compiled_reference():
    """oracle/ipm_numpy.EcosExact (equilibration, stretched KKT, regularisation, refinement, safeguards) against the
    golden vectors of the compiled reference: same iteration counts, x/y/s to 1e-9, z to 1e-6."""
    from oracle.ipm_numpy import EcosExact
    g = _golden()
    fam = families.portfolio_socp()
    c0, b0, h = fam.canon_data('c'), fam.canon_data('b'), fam.canon_data('h')
    E = EcosExact(fam.canon_matrix('A'), fam.canon_matrix('G'), 601, [12, 102])
    for k in (0, 3):
        c = c0.copy(); c[:100] = -g['param_a'][k]
        b = b0.copy(); b[11:111] = -g['param_w_prev'][k]
        r = E.solve(c, b, h)
        assert r['exitflag'] == 0 and r['iter'] == g['iter'][k]
        assert _rel(r['x'], g['x'][k]) < 1e-8 and _rel(r['y'], g['y'][k]) < 1e-8 and _rel(r['s'], g['s'][k]) < 1e-8
        assert _rel(r['z'], g['z'][k]) < 1e-6


def _build_emu(setup, d):
    with open(os.path.join(d, 'cpg_ipm_family.h'), 'w') as f:
        f.write(codegen_ipm.family_header(setup))
    out = os.path.join(d, 'libipm_emu.so')
    r = subprocess.run(['g++', '-O2', '-std=c++17', '-shared', '-fPIC', '-I', d, '-I', os.path.join(ROOT, 'cvxpygen_b200', 'csrc'),
                        os.path.join(ROOT, 'tests', 'emu', 'ipm_emu.cpp'), '-o', out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return C.CDLL(out)


@pytest.fixture(scope='module')
def emu(setup, tmp_path_factory):
    return _build_emu(setup, str(tmp_path_factory.mktemp('ipm_emu')))


def _emu_solve(lib, setup, params, maxit=100):
    D = setup.defines
    B = params.shape[0]
    out = dict(prim=np.zeros((B, D['NPRIM'])), dual=np.zeros((B, D['NDUAL'])), x=np.zeros((B, D['N'])), y=np.zeros((B, D['P'])),
               z=np.zeros((B, D['M'])), s=np.zeros((B, D['M'])), obj=np.zeros(B), iter=np.zeros(B, np.int32),
               status=np.zeros(B, np.int32), pres=np.zeros(B), dres=np.zeros(B))
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    params = np.ascontiguousarray(params, dtype=np.float64)
    lib.ipm_emu_solve(setup.smem_blob, setup.gmem_blob, B, P(params), *[P(out[k]) for k in
                      ('prim', 'dual', 'x', 'y', 'z', 's', 'obj', 'iter', 'status', 'pres', 'dres')], maxit)
    return out


def test_kernel_phase_logic_on_host_matches_golden(emu, setup):
    """The kernel source compiled as host C++ (phases serialised) against the compiled reference's golden vectors."""
    g = _golden()
    B = 8
    out = _emu_solve(emu, setup, np.c_[g['param_a'][:B], g['param_w_prev'][:B]])
    assert np.array_equal(out['iter'], g['iter'][:B]) and (out['status'] == 0).all()
    for k in range(B):
        assert _rel(out['x'][k], g['x'][k]) < RTOL_PRIMAL and _rel(out['y'][k], g['y'][k]) < RTOL_PRIMAL
        assert _rel(out['s'][k], g['s'][k]) < RTOL_PRIMAL and _rel(out['z'][k], g['z'][k]) < RTOL_DUAL
    assert np.allclose(out['obj'], -g['pcost'][:B], rtol=0, atol=1e-9)            # maximisation: obj_val = -pcost
    fam = setup.family
    assert np.array_equal(out['prim'][:, :100], out['x'][:, fam.variables[0].indices])
    zl1 = fam.duals[2].indices[0]
    assert np.array_equal(out['dual'][:, 11], out['z'][:, zl1])


def test_kernel_phase_logic_maxit_and_best_iterate(emu, setup):
    """maxit = 5: ECOS returns the better of current / best iterate with exit flag -1 (or an inaccurate flag)."""
    g = _golden()
    out = _emu_solve(emu, setup, np.c_[g['param_a'][:2], g['param_w_prev'][:2]], maxit=5)
    assert (out['iter'] == 5).all() and np.isin(out['status'], (-1, 10)).all()
    if ref_ecos.available():
        fam = setup.family
        c0, b0 = fam.canon_data('c'), fam.canon_data('b')
        r = ref_ecos.RefECOS(c0, fam.canon_matrix('A'), b0, fam.canon_matrix('G'), fam.canon_data('h'), 601, [12, 102], maxit=5)
        Cb = np.tile(c0, (2, 1)); Cb[:, :100] = -g['param_a'][:2]
        Bb = np.tile(b0, (2, 1)); Bb[:, 11:111] = -g['param_w_prev'][:2]
        ref = r.solve_batch(c=Cb, b=Bb)
        assert np.array_equal(ref['exitflag'], out['status']) and np.array_equal(ref['iter'], out['iter'])
        assert _rel(out['x'], ref['x']) < 1e-7 and _rel(out['z'], ref['z']) < 1e-6


GENERIC = {'random_socp_30_8_20_3x5x4': lambda: families.random_socp(30, 8, 20, (3, 5, 4), seed=5),
           'random_socp_20_5_30_lp': lambda: families.random_socp(20, 5, 30, (), seed=6),
           'random_socp_12_0_10_6': lambda: families.random_socp(12, 0, 10, (6,), seed=7)}


def _check_against(g, out_x, out_y, out_z, out_s, status, iters, obj, sl=slice(None)):
    """exit flags and iteration counts identical; optimal instances to the parity bar; certificates by direction"""
    flag = g['exitflag'][sl]
    assert np.array_equal(status, flag) and np.array_equal(iters, g['iter'][sl])
    assert set(np.unique(flag)) == {0, 1, 2}
    for k in np.nonzero(flag == 0)[0]:
        assert _rel(out_x[k], g['x'][sl][k]) < RTOL_PRIMAL and _rel(out_s[k], g['s'][sl][k]) < RTOL_PRIMAL, k
        assert _rel(out_z[k], g['z'][sl][k]) < RTOL_DUAL, k
        if out_y.shape[1]:
            assert _rel(out_y[k], g['y'][sl][k]) < RTOL_DUAL, k
        assert abs(obj[k] - g['pcost'][sl][k]) < 1e-7 * max(1.0, abs(g['pcost'][sl][k]))
    for k in np.nonzero(flag == 1)[0]:          # certificate of primal infeasibility: (y, z) (ecos.c:1219-1233)
        assert _rel(out_z[k], g['z'][sl][k]) < 1e-4, k
    for k in np.nonzero(flag == 2)[0]:          # certificate of unboundedness: x
        assert _rel(out_x[k], g['x'][sl][k]) < 1e-4, k


@pytest.mark.parametrize('name', list(GENERIC))
def test_generic_conic_families_on_host_match_golden(name, tmp_path):
    """Three cones + equalities / a pure LP (no second-order cone) / no equalities: the same kernel source and table
    generator, host build, against the compiled reference's golden vectors incl. primal- and dual-infeasible instances."""
    fam = GENERIC[name]()
    st = ss.setup_socp_family(fam)
    D = st.defines
    assert D['NT'] <= ss.MAX_TAIL and (D['NSOC'], D['P']) == (len(fam.cone_dims['q']), fam.n_eq)
    lib = _build_emu(st, str(tmp_path))
    g = np.load(os.path.join(GOLDEN, f'socp_{name}.npz'))
    B = 32
    P = np.concatenate([g['param_' + p.name][:B] for p in fam.params], axis=1)
    out = _emu_solve(lib, st, P)
    _check_against(g, out['x'], out['y'][:, :fam.n_eq], out['z'], out['s'], out['status'], out['iter'], out['obj'], slice(0, B))


def _network_inputs(g, fam):
    names = standard.STANDARD['network_lp_50_10'][1]
    return np.concatenate([g['param_' + nm] for nm in names], axis=1), names


def test_reference_network_lp_on_host_matches_golden(tmp_path):
    """The reference's own LP test problem (tests/test_E2E_LP.py:15-36, network flow n = 50, m = 10, solved there with
    ECOS): a maximisation with no equalities and no second-order cone, four batched vector parameters, the routing matrix
    shared.  Kernel phase logic on the host against the compiled ECOS golden vectors."""
    fam = families.network_lp(50, 10)
    g = np.load(os.path.join(GOLDEN, 'socp_network_lp_50_10.npz'))
    P, names = _network_inputs(g, fam)
    st = ss.setup_socp_family(fam, names)
    assert (st.defines['NSOC'], st.defines['P'], st.defines['IS_MAX']) == (0, 0, 1)
    lib = _build_emu(st, str(tmp_path))
    B = 24
    out = _emu_solve(lib, st, P[:B])
    assert np.array_equal(out['status'], g['exitflag'][:B]) and np.array_equal(out['iter'], g['iter'][:B])
    for k in range(B):
        assert _rel(out['x'][k], g['x'][k]) < RTOL_PRIMAL and _rel(out['z'][k], g['z'][k]) < RTOL_DUAL
    assert np.allclose(out['obj'], -g['pcost'][:B], rtol=1e-9)                 # maximise w'f: obj_val = -pcost
    assert np.allclose(out['obj'], (g['param_w'][:B] * out['x']).sum(1), rtol=1e-7)
    # user-level gathers: f = x, d0 / d1 / d2 = capacity / lower / upper multipliers
    assert np.array_equal(out['prim'][:, :50], out['x']) and np.array_equal(out['dual'][:, :110], out['z'])


def test_reference_adp_socp_on_host_matches_golden(tmp_path):
    """The reference's SOCP test problem (tests/test_E2E_SOCP.py:15-63, one ADP step): two squared norms as rotated cones and
    two input-norm bounds -- four second-order cones, NO LP cone, no equalities -- with `f` per instance."""
    fam = families.adp_socp()
    g = np.load(os.path.join(GOLDEN, 'socp_adp_socp_6_3.npz'))
    st = ss.setup_socp_family(fam, ['f'])
    assert (st.defines['NSOC'], st.defines['P'], st.defines['L']) == (4, 0, 0)
    lib = _build_emu(st, str(tmp_path))
    B = 24
    out = _emu_solve(lib, st, g['param_f'][:B])
    assert np.array_equal(out['status'], g['exitflag'][:B]) and np.array_equal(out['iter'], g['iter'][:B])
    for k in range(B):
        assert _rel(out['x'][k], g['x'][k]) < 1e-6 and _rel(out['s'][k], g['s'][k]) < 1e-6
    assert np.allclose(out['obj'], g['pcost'][:B], rtol=1e-7)
    # the user-level problem: u = first six canonical variables (2 x 3, Fortran order); objective = the two squared norms
    u0 = out['x'][:, 0:6:2]
    G = fam.param('G').default.reshape(6, 3, order='F'); Rs = fam.param('Rsqrt').default
    obj = ((g['param_f'][:B] + u0 @ G.T) ** 2).sum(1) + ((Rs * u0) ** 2).sum(1)
    assert np.allclose(obj, out['obj'], rtol=1e-5, atol=1e-7)
    assert np.linalg.norm(u0, axis=1).max() <= 0.1 + 1e-7 and np.linalg.norm(out['x'][:, 1:6:2], axis=1).max() <= 0.1 + 1e-7


@pytest.mark.gpu
def test_gpu_reference_adp_socp():
    g = np.load(os.path.join(GOLDEN, 'socp_adp_socp_6_3.npz'))
    m = standard.load('adp_socp_6_3')
    r = m.solve_batch({'f': g['param_f']}, return_canonical=True)
    assert np.array_equal(r.cpg_info.status, g['exitflag']) and np.array_equal(r.cpg_info.iter, g['iter'])
    assert _rel(r.sol_x, g['x']) < 1e-6 and _rel(r.sol_s, g['s']) < 1e-6
    assert np.allclose(r.cpg_info.obj_val, g['pcost'], rtol=1e-7)
    assert r.cpg_prim['u'].shape == (len(g['iter']), 2, 3) and np.array_equal(r.cpg_prim['u'][:, 0, :], r.sol_x[:, 0:6:2])


@pytest.mark.gpu
def test_gpu_reference_network_lp():
    fam = families.network_lp(50, 10)
    g = np.load(os.path.join(GOLDEN, 'socp_network_lp_50_10.npz'))
    m = standard.load('network_lp_50_10')
    names = standard.STANDARD['network_lp_50_10'][1]
    r = m.solve_batch({nm: g['param_' + nm] for nm in names}, return_canonical=True)
    assert np.array_equal(r.cpg_info.status, g['exitflag']) and np.array_equal(r.cpg_info.iter, g['iter'])
    assert _rel(r.sol_x, g['x']) < RTOL_PRIMAL and _rel(r.sol_z, g['z']) < RTOL_DUAL
    assert np.allclose(r.cpg_info.obj_val, -g['pcost'], rtol=1e-9)
    assert np.array_equal(r.cpg_prim['f'], r.sol_x) and np.array_equal(r.cpg_dual['d0'], r.sol_z[:, :10])
    # a fresh, larger batch: feasibility and optimality of every instance's own LP
    B = 5000
    par = families.network_lp_batch(fam, B, seed=77)
    r2 = m.solve_batch(par, return_canonical=True)
    assert (r2.cpg_info.status == 0).all()
    R = fam.param('R').default.reshape(10, 50, order='F')
    f = r2.cpg_prim['f']
    assert (f @ R.T - par['c']).max() < 1e-7 and (par['f_min'] - f).max() < 1e-7 and (f - par['f_max']).max() < 1e-7
    gap = np.abs((par['w'] * f).sum(1) - (-(r2.sol_z[:, :10] * par['c']).sum(1) + (r2.sol_z[:, 10:60] * par['f_min']).sum(1)
                                          - (r2.sol_z[:, 60:] * par['f_max']).sum(1)) * -1)
    assert (gap < 1e-6 * np.maximum(1.0, np.abs(r2.cpg_info.obj_val))).all()     # strong duality: w'f = c'z0 - fmin'z1 + fmax'z2


def _parse_c_arrays(path):
    """{name: numpy array or scalar} of the `pfloat` / `idxint` definitions of an ECOS test header"""
    import re
    txt = open(path).read()
    out = {}
    for m_ in re.finditer(r'(pfloat|idxint)\s+(\w+)\[\d*\]\s*=\s*\{([^}]*)\}', txt):
        vals = [v for v in re.split(r'[\s,]+', m_.group(3).strip()) if v]
        out[m_.group(2)] = np.array([float(v) for v in vals]) if m_.group(1) == 'pfloat' else np.array([int(v) for v in vals])
    for m_ in re.finditer(r'(pfloat|idxint)\s+(\w+)\s*=\s*([-+0-9.eE]+)\s*;', txt):
        out[m_.group(2)] = float(m_.group(3)) if m_.group(1) == 'pfloat' else int(m_.group(3))
    return out


ECOS_UPDATE_KAT = os.path.join(os.environ.get('CPG_REFERENCE', '/root/reference'), 'cvxpygen', 'solvers', 'ecos', 'test', 'updateData',
                               'update_data.h')


@pytest.mark.skipif(not os.path.exists(ECOS_UPDATE_KAT), reason='reference tree not present')
def test_ecos_update_data_known_answers(tmp_path):
    """ECOS's OWN known-answer test for ECOS_updateData (ecos/test/updateData/update_data.h, run by ecostester.c): an LP with
    n = 20, p = 5, l = 40 solved for (c1, A1, b1, G1, h1) -> optimal value -36.250515, then for the second data set -> -20.011586.
    The header is parsed where it lies; compiled reference (when built), the numpy restatement and the interior-point kernel's
    phase logic (host build) must reproduce both values.  Matrices are shared parameters here (one family per data set, the
    role of update_shared_params); c, b, h are the per-instance vectors."""
    from cvxpygen_b200.ir import CanonFamily
    from oracle.ipm_numpy import EcosExact
    d = _parse_c_arrays(ECOS_UPDATE_KAT)
    n, m, p, l = d['udd_n'], d['udd_m'], d['udd_p'], d['udd_l']
    for k, want in (('1', d['udd_optval1']), ('2', d['udd_optval2'])):
        A = sp.csc_matrix((d[f'udd_A{k}pr'], d['udd_Air'], d['udd_Ajc']), shape=(p, n))
        G = sp.csc_matrix((d[f'udd_G{k}pr'], d['udd_Gir'], d['udd_Gjc']), shape=(m, n))
        c, b, h = d[f'udd_c{k}'], d[f'udd_b{k}'], d[f'udd_h{k}']
        if ref_ecos.available():
            r = ref_ecos.RefECOS(c, A, b, G, h, l, []).solve_batch()
            assert r['exitflag'][0] in (0, 10) and abs(r['pcost'][0] - want) < 1e-5
        e = EcosExact(A, G, l, []).solve(c, b, h)
        assert e['exitflag'] in (0, 10) and abs(e['pcost'] - want) < 1e-5
        fam = CanonFamily.from_canonical_conic(f'ecos_update_data_{k}', c, A, b, G, h, l)
        st = ss.setup_socp_family(fam, ['c', 'b', 'h'])
        os.makedirs(str(tmp_path / k), exist_ok=True)
        lib = _build_emu(st, str(tmp_path / k))
        out = _emu_solve(lib, st, np.concatenate([c, b, h])[None, :])
        assert out['status'][0] in (0, 10) and abs(out['obj'][0] - want) < 1e-5
        assert out['iter'][0] == e['iter'] and np.abs(out['x'][0] - e['x']).max() < 1e-7 * max(1.0, np.abs(e['x']).max())


ECOS_TESTS = os.path.join(os.environ.get('CPG_REFERENCE', '/root/reference'), 'cvxpygen', 'solvers', 'ecos', 'test')


@pytest.mark.skipif(not os.path.isdir(ECOS_TESTS), reason='reference tree not present')
@pytest.mark.parametrize('rel,prefix,flag', [('infeasibleProblems/infeasible1.h', '', 1), ('infeasibleProblems/infeasible2.h', '', 1),
                                             ('unboundedProblems/unboundedLP1.h', '', 2), ('LPnetlib/lp_afiro.h', 'lp_afiro_', 0)])
def test_ecos_own_test_problems(rel, prefix, flag, tmp_path):
    """Problems of ECOS's own test suite (ecos/test/*, run by ecostester.c) with the exit flag each of them asserts: primal
    infeasible (with and without equalities), unbounded LP, a netlib LP.  (unboundedMaxSqrt.h is left out: it sits on a
    numerical knife edge -- the vendored ECOS itself, compiled here, returns ECOS_NUMERICS after 11 iterations from a fresh
    ECOS_setup and ECOS_DINF after 12 once ECOS_updateData has re-equilibrated the very same data.)  Parsed where they lie; compiled reference (when built), numpy restatement and the kernel's phase logic (host build)."""
    from cvxpygen_b200.ir import CanonFamily
    from oracle.ipm_numpy import EcosExact
    d = _parse_c_arrays(os.path.join(ECOS_TESTS, rel))
    g = lambda k, default=None: d.get(prefix + k, default)
    n, m, p, l = g('n'), g('m'), g('p'), g('l')
    q = [int(v) for v in g('q', [])] if g('ncones', 0) else []
    G = sp.csc_matrix((g('Gpr'), g('Gir'), g('Gjc')), shape=(m, n))
    A = sp.csc_matrix((g('Apr'), g('Air'), g('Ajc')), shape=(p, n)) if p else sp.csc_matrix((0, n))
    c, h, b = g('c'), g('h'), (g('b') if p else np.zeros(0))
    e = EcosExact(A, G, l, q).solve(c, b, h)
    assert e['exitflag'] == flag
    if ref_ecos.available():
        r = ref_ecos.RefECOS(c, A, b, G, h, l, q).solve_batch()
        assert r['exitflag'][0] == flag and r['iter'][0] == e['iter']
    fam = CanonFamily.from_canonical_conic('ecos_kat', c, A, b, G, h, l, q)
    st = ss.setup_socp_family(fam, [pp.name for pp in fam.params])
    lib = _build_emu(st, str(tmp_path))
    out = _emu_solve(lib, st, np.concatenate([c, b, h])[None, :])
    assert out['status'][0] == flag and out['iter'][0] == e['iter']
    if flag == 0:
        assert abs(out['obj'][0] - e['pcost']) < 1e-7 * max(1.0, abs(e['pcost'])) and _rel(out['x'][0], e['x']) < 1e-6


def test_generated_directory_and_c_abi():
    d = standard.build(NAME)
    for f in ('cpg_solver.py', 'cpg_module.py', 'cpg_meta.json', 'libcpg_b200.so', 'c/include/cpg_b200_socp.h',
              'c/include/cpg_ipm_family.h', 'c/solver_code/ipm_kernel.cuh', 'c/solver_code/cpg_b200_ipm_module.cu',
              'c/src/cpg_ipm_blob.c'):
        assert os.path.exists(os.path.join(d, f)), f
    hdr = open(os.path.join(ROOT, 'include', 'cpg_b200_socp.h')).read()
    declared = set(re.findall(r'CPG_B200_FN\((\w+)\)\s*\(', hdr))
    assert {'cpg_b200_init', 'cpg_socp_solve_batch_device', 'cpg_socp_solve_batch_host', 'cpg_socp_dims'} <= declared
    lib = C.CDLL(os.path.join(d, 'libcpg_b200.so'))
    for sym in declared:
        assert hasattr(lib, sym), sym
    m = runtime.load(d)
    assert isinstance(m, runtime.SocpModule)
    assert (m.dims.n_var, m.dims.n_eq, m.dims.n_ineq, m.dims.n_lp, m.dims.n_soc) == (512, 111, 715, 601, 2)
    assert (m.dims.n_param, m.dims.n_prim, m.dims.n_dual) == (200, 210, 112)
    assert m.dims.smem_bytes <= 232448 - 1024
    assert (m.settings.maxit, m.settings.feastol, m.settings.reltol_inacc) == (100, 1e-8, 5e-5)
    with pytest.raises(AttributeError):
        m.set_solver_setting('eps_abs', 1e-3)
    with pytest.raises(AttributeError):
        m.pack_params({'nope': np.zeros(3)})


def test_wrong_solver_for_family_is_rejected(tmp_path):
    from cvxpygen_b200 import cpg
    with pytest.raises(ValueError):
        cpg.generate_code(families.portfolio_socp(), code_dir=str(tmp_path / 'x'), solver='ADMM-CUDA', wrapper=False)
    # the other way round is allowed, as in the reference (QPs run under ECOS too): the family is restated in conic form
    cpg.generate_code(families.mpc(2, 1, 2), code_dir=str(tmp_path / 'y'), solver='IPM-CUDA', wrapper=False)
    assert os.path.exists(str(tmp_path / 'y' / 'c' / 'include' / 'cpg_ipm_family.h'))
    with pytest.raises(ValueError):       # a conic family has no QP backward pass (extended DPP, cvxpygen/canonicalizer.py:338-345)
        cpg.generate_code(families.portfolio_socp(), code_dir=str(tmp_path / 'z'), solver='IPM-CUDA', gradient=True, wrapper=False)


# ---------------------------------------------------------------------------------------------------------------------
def _ref_batch(fam, a, wp, **kw):
    c0, b0 = fam.canon_data('c'), fam.canon_data('b')
    r = ref_ecos.RefECOS(c0, fam.canon_matrix('A'), b0, fam.canon_matrix('G'), fam.canon_data('h'),
                         fam.cone_dims['l'], fam.cone_dims['q'], **kw)
    B = a.shape[0]
    Cb = np.tile(c0, (B, 1)); Cb[:, :100] = -a
    Bb = np.tile(b0, (B, 1)); Bb[:, 11:111] = -wp
    return r.solve_batch(c=Cb, b=Bb)


@pytest.mark.gpu
def test_gpu_portfolio_matches_golden():
    g = _golden()
    m = standard.load(NAME)
    r = m.solve_batch({'a': g['param_a'], 'w_prev': g['param_w_prev']}, return_canonical=True)
    assert m.launch_count() == 1
    assert np.array_equal(r.cpg_info.iter, g['iter']) and (r.cpg_info.status == 0).all()
    for k in range(g['x'].shape[0]):
        assert _rel(r.sol_x[k], g['x'][k]) < RTOL_PRIMAL and _rel(r.sol_y[k], g['y'][k]) < RTOL_PRIMAL
        assert _rel(r.sol_s[k], g['s'][k]) < RTOL_PRIMAL and _rel(r.sol_z[k], g['z'][k]) < RTOL_DUAL
    assert np.allclose(r.cpg_info.obj_val, -g['pcost'], rtol=0, atol=1e-9)
    fam = families.portfolio_socp()
    assert np.array_equal(r.cpg_prim['w'], r.sol_x[:, fam.variables[0].indices])
    assert np.array_equal(r.cpg_dual['d3'], r.sol_y[:, fam.duals[3].indices])


@pytest.mark.gpu
@pytest.mark.skipif(not ref_ecos.available(), reason='oracle/_ref/libecos_ref.so not built')
def test_gpu_portfolio_matches_compiled_reference_on_fresh_batch():
    fam = families.portfolio_socp()
    rng = np.random.default_rng(7)
    B = 400
    a = rng.standard_normal((B, 100)) * rng.uniform(0.2, 2.0, (B, 1))
    wp = np.abs(1 / 100 + 0.02 * rng.standard_normal((B, 100)))
    ref = _ref_batch(fam, a, wp)
    m = standard.load(NAME)
    r = m.solve_batch({'a': a, 'w_prev': wp}, return_canonical=True)
    assert np.array_equal(r.cpg_info.status, ref['exitflag'])
    assert (np.abs(r.cpg_info.iter - ref['iter']) <= 1).all() and (r.cpg_info.iter == ref['iter']).mean() > 0.97
    same = r.cpg_info.iter == ref['iter']
    ez = np.array([_rel(r.sol_z[k], ref['z'][k]) for k in range(B)])
    for k in np.nonzero(same)[0]:
        assert _rel(r.sol_x[k], ref['x'][k]) < RTOL_PRIMAL, k
        assert _rel(r.sol_y[k], ref['y'][k]) < RTOL_PRIMAL, k
        assert _rel(r.sol_s[k], ref['s'][k]) < RTOL_PRIMAL, k
    # duals: the two rows  +-(w - w_prev)_i <= t_i  are both active when asset i is not traded, their multipliers are then
    # determined only up to a shift along (+1, -1) and an interior-point method fixes that shift to a few digits.  One
    # instance in a few hundred lands elsewhere on that face than the reference (measured: the entries move in exact
    # +d / -d pairs).  Bar: 1e-5 on at least 99 % of the batch, 1e-3 on all, and on EVERY instance the returned duals
    # satisfy dual feasibility and complementarity as tightly as the reference's own do.
    assert (ez[same] < RTOL_DUAL).mean() >= 0.99 and ez[same].max() < 1e-3, (ez.max(), (ez < RTOL_DUAL).mean())
    A, G = fam.canon_matrix('A').toarray(), fam.canon_matrix('G').toarray()
    Cb = np.tile(fam.canon_data('c'), (B, 1)); Cb[:, :100] = -a
    assert np.abs(r.sol_y @ A + r.sol_z @ G + Cb).max() < 1e-9
    assert np.abs((r.sol_s * r.sol_z).sum(1)).max() < 1e-7
    # an instance that stops one iteration earlier / later still agrees to the solver tolerance on the user variables
    assert _rel(r.cpg_prim['w'], ref['x'][:, :100]) < 1e-5
    assert np.allclose(r.cpg_info.obj_val, -ref['pcost'], rtol=0, atol=1e-7)


@pytest.mark.gpu
def test_gpu_portfolio_full_batch_properties():
    """BASELINE config 3 size (batch 50k): optimality conditions of the ORIGINAL (un-equilibrated) problem on every
    instance, device-buffer entry point."""
    import torch
    fam = families.portfolio_socp()
    B = 50000
    rng = np.random.default_rng(11)
    a = rng.standard_normal((B, 100)); wp = np.abs(1 / 100 + 0.01 * rng.standard_normal((B, 100)))
    m = standard.load(NAME)
    P = torch.from_numpy(np.ascontiguousarray(np.c_[a, wp])).cuda()
    out = m.solve_batch_device(P, return_canonical=True)
    torch.cuda.synchronize()
    st = out.status.cpu().numpy()
    # a handful of instances per 10^5 end "close to optimal" (exit flag 10): ECOS's safeguard |pres| > 500 |pres_prev|
    # (ecos.c:1180-1200) compares two residuals that are both rounding noise (1e-11 vs 1e-14) in the last iteration, so which
    # side of 500 the ratio falls on depends on the last bits -- the reference does the same on other instances.  They must
    # be rare, the reference must call them solved as well, and the returned (best previous) iterate must agree with the
    # reference's answer to the reduced tolerance; the optimality conditions below are asserted on the rest
    odd = np.nonzero(st != 0)[0]
    assert odd.size <= B // 10000 and np.isin(st[odd], (0, 10)).all()
    if odd.size and ref_ecos.available():
        ref = _ref_batch(fam, a[odd], wp[odd])
        assert np.isin(ref['exitflag'], (0, 10)).all()
        assert _rel(out.sol_x.cpu().numpy()[odd], ref['x']) < 1e-4
    ok = st == 0
    x, y, z, s = (t.cpu().numpy()[ok] for t in (out.sol_x, out.sol_y, out.sol_z, out.sol_s))
    a, wp, B_all, B = a[ok], wp[ok], B, int(ok.sum())
    A, G = fam.canon_matrix('A'), fam.canon_matrix('G')
    c0, b0, h = fam.canon_data('c'), fam.canon_data('b'), fam.canon_data('h')
    Cb = np.tile(c0, (B, 1)); Cb[:, :100] = -a
    Bb = np.tile(b0, (B, 1)); Bb[:, 11:111] = -wp
    assert np.abs(x @ A.T.toarray() - Bb).max() < 1e-6
    assert np.abs(h[None, :] - x @ G.T.toarray() - s).max() < 1e-6
    assert np.abs(y @ A.toarray() + z @ G.toarray() + Cb).max() < 1e-6
    assert (s[:, :601] > -1e-9).all() and (z[:, :601] > -1e-9).all()
    assert (s[:, 601] >= np.linalg.norm(s[:, 602:613], axis=1) - 1e-8).all()
    assert (z[:, 613] >= np.linalg.norm(z[:, 614:], axis=1) - 1e-8).all()
    assert np.abs((s * z).sum(1)).max() < 1e-5
    w = x[:, :100]
    assert np.abs(w.sum(1) - 1).max() < 1e-7 and (np.abs(w).sum(1) <= 1.6 + 1e-6).all()
    # the same batch twice gives the same answer to rounding (accumulation order of shared-memory atomics may differ)
    out2 = m.solve_batch_device(P, return_canonical=True)
    torch.cuda.synchronize()
    assert (out2.iter == out.iter).all() and _rel(out2.sol_x.cpu().numpy()[ok], x) < 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize('name', [n for n in standard.SOCP_NAMES if n.startswith('random_socp')])
def test_gpu_generic_conic_families(name):
    """The generic conic families through the C ABI: golden vectors (exit flags 0 / 1 / 2, iteration counts, solutions and
    certificates) and a fresh batch against the compiled reference."""
    from helpers import conic_batch
    fam = standard.STANDARD[name][0]()
    g = np.load(os.path.join(GOLDEN, f'socp_{name}.npz'))
    m = standard.load(name)
    par = {p.name: g['param_' + p.name] for p in fam.params}
    r = m.solve_batch(par, return_canonical=True)
    _check_against(g, r.sol_x, r.sol_y, r.sol_z, r.sol_s, r.cpg_info.status, r.cpg_info.iter, r.cpg_info.obj_val)
    assert np.array_equal(r.cpg_prim['x'], r.sol_x)
    if ref_ecos.available():
        par2, kind = conic_batch(fam, 512, seed=7)
        ref = ref_ecos.RefECOS(fam.canon_data('c'), fam.canon_matrix('A'), fam.canon_data('b'), fam.canon_matrix('G'),
                               fam.canon_data('h'), fam.cone_dims['l'], fam.cone_dims['q']).solve_batch(c=par2['c'], h=par2['h'], b=par2.get('b'))
        r2 = m.solve_batch(par2, return_canonical=True)
        assert np.array_equal(r2.cpg_info.status, ref['exitflag'])
        assert (np.abs(r2.cpg_info.iter - ref['iter']) <= 1).all() and (r2.cpg_info.iter == ref['iter']).mean() > 0.97
        ok = (ref['exitflag'] == 0) & (r2.cpg_info.iter == ref['iter'])
        assert _rel(r2.sol_x[ok], ref['x'][ok]) < 1e-6 and np.allclose(r2.cpg_info.obj_val[ok], ref['pcost'][ok], rtol=1e-7, atol=1e-8)


@pytest.mark.gpu
@pytest.mark.skipif(not ref_ecos.available(), reason='oracle/_ref/libecos_ref.so not built')
def test_gpu_shared_parameter_update(tmp_path):
    """A user parameter that is NOT batched is shared by every instance: changing it re-runs the offline setup on the host
    and re-uploads both constant images (cpg_socp_load_constants) -- the role of ECOS_updateData for shared data."""
    from cvxpygen_b200 import cpg
    from helpers import conic_batch
    fam = families.random_socp(20, 5, 30, (4,), seed=8)
    d = str(tmp_path / 'conic_shared')
    cpg.generate_code(fam, code_dir=d, solver='IPM-CUDA', batch_params=['c', 'h'], wrapper=True)
    m = runtime.load(d)
    par, kind = conic_batch(fam, 96, seed=21)
    ref = ref_ecos.RefECOS(fam.canon_data('c'), fam.canon_matrix('A'), fam.canon_data('b'), fam.canon_matrix('G'),
                           fam.canon_data('h'), fam.cone_dims['l'], fam.cone_dims['q'])

    def check(b):
        r = m.solve_batch({'c': par['c'], 'h': par['h']}, return_canonical=True)
        o = ref.solve_batch(c=par['c'], h=par['h'], b=np.tile(b, (96, 1)))
        assert np.array_equal(r.cpg_info.status, o['exitflag']) and (np.abs(r.cpg_info.iter - o['iter']) <= 1).all()
        ok = (o['exitflag'] == 0) & (r.cpg_info.iter == o['iter'])
        assert ok.sum() > 40 and _rel(r.sol_x[ok], o['x'][ok]) < 1e-6
        return r
    r0 = check(fam.param('b').default)
    b_new = fam.param('b').default + 0.05 * np.random.default_rng(3).standard_normal(fam.n_eq)
    m.update_shared_params({'b': b_new})
    r1 = check(b_new)
    assert _rel(r1.sol_x, r0.sol_x) > 1e-4                      # the update is visible in the solutions
    with pytest.raises(ValueError):
        m.update_shared_params({'c': par['c'][0]})                # batched parameters go through solve_batch
    with pytest.raises(AttributeError):
        m.update_shared_params({'nope': 1.0})


# ---------------------------------------------------------------------------------------------------------------------
# per-instance MATRIX parameters on the conic path (VERDICT r1 missing item 4): F enters A, d_sqrt enters G
def _portfolio_mat_batch(fam, B, seed=3):
    rng = np.random.default_rng(seed)
    a = fam.param('a').default[None, :] + 0.3 * rng.standard_normal((B, fam.param('a').size))
    wp = np.abs(rng.standard_normal((B, fam.param('w_prev').size))); wp /= wp.sum(1, keepdims=True)
    F = fam.param('F').default[None, :] + 0.25 * rng.standard_normal((B, fam.param('F').size))       # every entry moves (zeros too)
    d = fam.param('d_sqrt').default[None, :] * rng.uniform(0.5, 1.5, (B, fam.param('d_sqrt').size))
    return dict(a=a, w_prev=wp, F=F, d_sqrt=d)


def _conic_reference_mat(fam, params, B, **kw):
    """Compiled ECOS driven like the reference's generated code when G / A are outdated: raw values + ECOS_updateData per instance."""
    th = np.tile(fam.theta_default(), (B, 1))
    for pn, v in params.items():
        p = fam.param(pn)
        th[:, p.col:p.col + p.size] = v
    data = {k: np.asarray(th @ fam.maps[k].T.toarray()) for k in ('c', 'b', 'h', 'A', 'G')}
    r = ref_ecos.RefECOS(fam.canon_data('c'), fam.canon_matrix('A'), fam.canon_data('b'), fam.canon_matrix('G'), fam.canon_data('h'),
                         fam.cone_dims['l'], fam.cone_dims['q'], **kw)
    return r.solve_batch(c=data['c'], b=data['b'], h=data['h'], G=data['G'], A=data['A'])


@pytest.mark.skipif(not ref_ecos.available(), reason='oracle/_ref/libecos_ref.so not built')
def test_conic_matrix_parameters_on_host_match_compiled_reference(tmp_path):
    """F (in A) and d_sqrt (in G) per instance: the kernel canonicalises the entries, re-equilibrates them (three Ruiz passes with
    exact maxima), rebuilds K from them and un-scales with its own scalings -- against ECOS_updateData + ECOS_solve of the compiled
    reference on the same raw values: identical exit flags and iteration counts."""
    fam = families.portfolio_socp(20, 4, matrix_params=True)
    names = ['a', 'w_prev', 'F', 'd_sqrt']
    st = ss.setup_socp_family(fam, names)
    assert st.defines['MATPAR'] == 1 and st.defines['NEMAP'] == 20 * 4 + 20
    lib = _build_emu(st, str(tmp_path))
    B = 12
    params = _portfolio_mat_batch(fam, B)
    P = np.concatenate([params[nm] for nm in names], axis=1)
    out = _emu_solve(lib, st, P)
    ref = _conic_reference_mat(fam, params, B)
    assert (ref['exitflag'] == 0).all()
    assert np.array_equal(out['status'], ref['exitflag']) and np.array_equal(out['iter'], ref['iter'])
    for k in range(B):
        assert _rel(out['x'][k], ref['x'][k]) < RTOL_PRIMAL and _rel(out['s'][k], ref['s'][k]) < RTOL_PRIMAL, k
        assert _rel(out['y'][k], ref['y'][k]) < RTOL_DUAL and _rel(out['z'][k], ref['z'][k]) < RTOL_DUAL, k
    # default matrix values reproduce the shared-matrix family's solutions
    fam0 = families.portfolio_socp(20, 4)
    st0 = ss.setup_socp_family(fam0, ['a', 'w_prev'])
    lib0 = _build_emu(st0, str(tmp_path / 'shared') if os.makedirs(str(tmp_path / 'shared'), exist_ok=True) is None else None)
    P0 = np.concatenate([params['a'], params['w_prev']], axis=1)
    Pd = np.concatenate([params['a'], params['w_prev'], np.tile(fam.param('F').default, (B, 1)), np.tile(fam.param('d_sqrt').default, (B, 1))], axis=1)
    o0, od = _emu_solve(lib0, st0, P0), _emu_solve(lib, st, Pd)
    assert np.array_equal(o0['iter'], od['iter']) and np.abs(o0['x'] - od['x']).max() < 1e-7


@pytest.mark.gpu
@pytest.mark.parametrize('name,B', [('portfolio_socp_mat_20_4', 256), ('portfolio_socp_mat_100_10', 96)])
def test_gpu_conic_matrix_parameters(name, B):
    """Per-instance G / A values on the GPU (ipm_kernel, IPM_MATPAR) against the compiled reference driven through ECOS_updateData
    with the same raw values: identical exit flags and iteration counts, user-level primal / dual at the parity bar; bitwise
    reproducible (the equilibration's maxima are exact, every sum has a fixed order)."""
    import time
    fam = standard.STANDARD[name][0]()
    names = standard.STANDARD[name][1]
    mod = standard.load(name)
    params = _portfolio_mat_batch(fam, B, seed=11)
    res = mod.solve_batch(params, return_canonical=True)
    ref = _conic_reference_mat(fam, params, B)
    assert (ref['exitflag'] == 0).all()
    assert np.array_equal(res.cpg_info.status, ref['exitflag']) and np.array_equal(res.cpg_info.iter, ref['iter'])
    for k in range(B):
        assert _rel(res.sol_x[k], ref['x'][k]) < RTOL_PRIMAL and _rel(res.sol_s[k], ref['s'][k]) < RTOL_PRIMAL, k
        assert _rel(res.sol_y[k], ref['y'][k]) < RTOL_DUAL and _rel(res.sol_z[k], ref['z'][k]) < RTOL_DUAL, k
    prim_ref = np.concatenate([ref['x'][:, v.indices] for v in fam.variables], axis=1)
    assert np.abs(res.prim - prim_ref).max() < 1e-6 * max(1.0, np.abs(prim_ref).max())
    again = mod.solve_batch(params, return_canonical=True)
    assert np.array_equal(again.sol_x, res.sol_x) and np.array_equal(again.sol_z, res.sol_z)
    if name.endswith('100_10'):
        Bt = 4096
        pt = _portfolio_mat_batch(fam, Bt, seed=12)
        mod.solve_batch(pt)
        t0 = time.perf_counter(); r2 = mod.solve_batch(pt); dt = time.perf_counter() - t0
        print(f'\n{name}: {Bt} instances in {dt * 1e3:.1f} ms (host call) = {Bt / dt:.0f} inst/s, mean iter {r2.cpg_info.iter.mean():.2f}, '
              f'optimal {float((r2.cpg_info.status == 0).mean()):.4f}')
