"""Host-side (generation-time) pipeline: equilibration, ordering + LDL', warp schedule, refactor tables, blob."""
import os
import struct

import numpy as np
import pytest
import scipy.sparse as sp

from cvxpygen_b200 import families, standard
from cvxpygen_b200.offline import kkt, refactor, schedule
from cvxpygen_b200.offline.blob import HEADER_FIELDS, MAGIC
from cvxpygen_b200.offline.qp_setup import setup_qp_family
from helpers import GOLDEN

NAMES = list(standard.QP_NAMES)


@pytest.fixture(scope='module')
def setups():
    return {n: setup_qp_family(standard.STANDARD[n][0](), standard.STANDARD[n][1]) for n in NAMES}


@pytest.mark.parametrize('name', NAMES)
def test_equilibration_bit_exact_vs_reference(setups, name):
    """D, E, c must equal the compiled reference's scale_data output bit for bit (golden from oracle/_ref)."""
    g = np.load(os.path.join(GOLDEN, f'{name}.npz'))
    st = setups[name]
    assert np.array_equal(st.D, g['D']) and np.array_equal(st.E, g['E']) and st.c == float(g['c'])


@pytest.mark.parametrize('name', NAMES)
def test_factor_and_schedule_solve_the_kkt_system(setups, name):
    st = setups[name]
    K = kkt.assemble_kkt(st.P_scaled, st.A_scaled, st.sigma, kkt.rho_vector(st.ctype, st.rho)).toarray()
    F = st.factor
    Lf = F.L + np.eye(K.shape[0])
    assert np.abs(Lf @ np.diag(F.D) @ Lf.T - K[np.ix_(F.perm, F.perm)]).max() < 1e-9 * np.abs(K).max()
    assert (np.diff(F.level) >= 0).all() and F.n_pos == st.n
    assert not (F.L != 0)[~F.Lpattern].any()
    b = np.random.default_rng(0).standard_normal((7, K.shape[0]))
    w = st.schedule.apply(b[:, F.perm])
    x = np.empty_like(w); x[:, F.perm] = w
    xs = np.linalg.solve(K, b.T).T
    assert np.abs(x - xs).max() < 1e-9 * np.abs(xs).max()


@pytest.mark.parametrize('name', NAMES)
def test_schedule_hazard_rules(setups, name):
    """A tile may only read rows that earlier tiles of its own group have not yet overwritten: forward tiles
    read positions <= their largest row, backward tiles read positions >= their smallest row."""
    S = setups[name].schedule
    for it, t in enumerate(S.tiles):
        used = t.cols[t.vals != 0]
        if it < S.n_fwd_tiles:
            assert used.max(initial=0) <= t.rows.max()
        elif it >= S.n_fwd_tiles + S.n_trailing_tiles:
            assert used.min(initial=65535) >= t.rows.min()
        assert t.r_pad in (1, 2, 4, 8, 16, 32) and len(t.rows) <= t.r_pad


@pytest.mark.parametrize('name', NAMES)
@pytest.mark.parametrize('rho', [0.1, 0.0071, 12.5])
def test_refactor_tables_emulation(setups, name, rho):
    st = setups[name]
    rv = kkt.rho_vector(st.ctype, rho)
    K = kkt.assemble_kkt(st.P_scaled, st.A_scaled, st.sigma, rv).toarray()
    T = st.refactor
    S = refactor.emulate_factor(T, rv)
    b = np.random.default_rng(1).standard_normal(K.shape[0])
    w = refactor.emulate_solve(T, S, b[st.factor.perm])
    x = np.empty_like(w); x[st.factor.perm] = w
    xs = np.linalg.solve(K, b)
    assert np.abs(x - xs).max() < 1e-8 * np.abs(xs).max()
    assert T.n_slots < 65536 and (T.ops[:, 0] < T.n_slots).all()


def test_blob_header_roundtrip(setups):
    st = setups['mpc_12_4_10']
    fmt = '<' + ''.join('i' if t == 'int' else 'd' for t, _ in HEADER_FIELDS)
    vals = dict(zip([n for _, n in HEADER_FIELDS], struct.unpack(fmt, st.blob[:struct.calcsize(fmt)])))
    assert vals['magic'] == MAGIC and vals['total_bytes'] == len(st.blob) and len(st.blob) % 16 == 0
    assert (vals['n'], vals['m'], vals['npb']) == (172, 172, 12)
    assert vals['n_tiles'] == len(st.schedule.tiles) and vals['rho'] == 0.1 and vals['sigma'] == 1e-6
    assert vals['off_i32'] % 16 == 0 and vals['off_f64'] % 16 == 0 and vals['off_u16'] % 16 == 0
    assert vals['i_tiles'] % 4 == 0          # int4 loads of tile headers
    assert len(st.blob) < 200 * 1024         # must fit shared memory next to the work vectors


def test_minimum_degree_and_levels_on_arrow_matrix():
    n = 12
    M = sp.lil_matrix((n, n)); M.setdiag(1.0); M[0, :] = 1; M[:, 0] = 1
    order = kkt.minimum_degree_order(sp.csr_matrix(M))
    assert order[-1] == 0 or order[-2] == 0      # the hub is eliminated (next to) last: no fill
    struct_, parent = kkt._symbolic(sp.csr_matrix(M), order)
    assert sum(len(s) for s in struct_) == n - 1


def test_matrix_parameter_as_batched_selects_the_matrix_path():
    fam = families.nonneg_ls(3, 2)
    st = setup_qp_family(fam, ['A'])                 # row f2: per-instance matrix parameters are generated
    assert st.mat_params == ['A'] and len(st.mat_blob) > 0
    assert len(setup_qp_family(fam, ['b']).mat_blob) == 0
    with pytest.raises(AttributeError):
        setup_qp_family(fam, ['nope'])


def test_constraint_types_follow_reference_thresholds():
    l = np.array([-1e30, 0.0, 1.0, -1e27]); u = np.array([1e30, 0.0, 1.00005, 5.0])
    assert kkt.constraint_types(l, u).tolist() == [-1, 1, 1, 0]
    assert np.allclose(kkt.rho_vector(np.array([-1, 0, 1]), 0.1), [1e-6, 0.1, 100.0])
