"""cvxpy-free builders of the problem families named by BASELINE.json.

cvxpy is not installable in the build container, so the families the benchmark
and the tests need are canonicalised by hand here into the same IR
(``cvxpygen_b200.ir.CanonFamily``) that the reference's ``Canonicalizer`` would
feed the solver plugin with (reference: cvxpygen/canonicalizer.py:86-122).
The hand-derived forms are the compact ones of SURVEY Appendix D: they have fewer
auxiliary variables than cvxpy's own canonicalisation, but the user-level
variables and constraint duals coincide at the optimum.  Row signs follow
cvxpy's convention (``lhs - rhs == 0`` for equalities, ``expr <= 0`` for
inequalities) so that dual signs match the reference's.

Families:
  nonneg_ls(m, n)      README example (reference: examples/main.py:16-26)
  mpc(nx, nu, N)       MPC QP (reference: tests/test_E2E_QP.py:44-73,127-146; BASELINE config 2)
  mpc_reference(H)     the reference's test MPC with all six parameters (diag / sparse matrix parameters)
  mpc_ltv(nx, nu, N)   the same with the dynamics and stage costs as (batchable) matrix parameters
"""
from typing import Optional

import numpy as np
import scipy.sparse as sp

from .ir import CanonFamily, UserParam, UserVar, UserDual

INF = 1e30  # OSQP_INFTY (reference: osqp_sources/include/constants.h:100; cvxpygen/utils.py:213-228 replace_inf)


class _MapBuilder:
    """Collects (entry, theta-column, coefficient) triplets of one affine map."""

    def __init__(self, n_rows, n_theta):
        self.n_rows, self.n_theta = n_rows, n_theta
        self.r, self.c, self.v = [], [], []

    def add(self, row, col, val):
        self.r.append(row); self.c.append(col); self.v.append(val)

    def const(self, row, val):
        self.add(row, self.n_theta - 1, val)

    def csr(self):
        M = sp.coo_matrix((self.v, (self.r, self.c)), shape=(self.n_rows, self.n_theta)).tocsr()
        M.sum_duplicates()
        return M


def _layout_params(specs):
    """specs: list of (name, shape, default_flat) -> list of UserParam with running columns."""
    params, col = [], 0
    for name, shape, default in specs:
        default = np.atleast_1d(np.asarray(default, dtype=float)).ravel()
        params.append(UserParam(name, tuple(shape), default.size, col, default))
        col += default.size
    return params


def _csc_pattern(M):
    M = sp.csc_matrix(M)
    M.sort_indices()
    return M.indices.astype(np.int32), M.indptr.astype(np.int32), M.shape


def default_mpc_dynamics(nx=12, nu=4, seed=0, dt=0.1):
    """The survey-probe plant: nx/2 positions + nx/2 damped
    velocities, nu force inputs.  Same construction and seed as BASELINE.md section 2."""
    h = nx // 2
    rs = np.random.RandomState(seed)
    Ad = np.eye(nx)
    Ad[:h, h:] = dt * np.eye(h)
    Ad[h:, h:] *= 0.98
    M = rs.randn(h, nu)
    Bd = np.zeros((nx, nu))
    Bd[h:, :] = dt * M
    Bd[:h, :] = 0.5 * dt * dt * M
    return Ad, Bd


def mpc(nx=12, nu=4, N=10, Ad=None, Bd=None, Q=None, QN=None, R=None, umax=1.0,
        x_init=None, d_const=0.0, name=None) -> CanonFamily:
    """MPC QP, compact sparse stacking:

        min  sum_{k<N} x_k'Q x_k + x_N'QN x_N + sum_{k<N} u_k'R u_k  (+ d_const)
        s.t. x_{k+1} = Ad x_k + Bd u_k,  x_0 = x_init,  |u_k|_inf <= umax

    canonical x = [X(:) ; U(:)]  (X is nx x (N+1), U is nu x N, Fortran order).
    P = 2*blkdiag(Q.., QN, R..) because the reference objective has no 1/2
    (reference: tests/test_E2E_QP.py:63-64).
    rows:  [0, nx)            x_0 = x_init                          (user constraint d2)
           [nx, (N+1)nx)      x_{k+1} - Ad x_k - Bd u_k = 0         (user constraint d0)
           [(N+1)nx, +N*nu)   -umax <= u <= umax                    (user constraint d1)
    Only ``x_init`` is a user parameter: it enters l and u of the first nx rows.
    """
    if Ad is None or Bd is None:
        Ad, Bd = default_mpc_dynamics(nx, nu)
    Q = np.eye(nx) if Q is None else np.asarray(Q, float)
    QN = Q if QN is None else np.asarray(QN, float)
    R = 0.1 * np.eye(nu) if R is None else np.asarray(R, float)
    if x_init is None:
        x_init = np.zeros(nx)
    nX, nU = (N + 1) * nx, N * nu
    n, n_eq, n_ineq = nX + nU, nX, nU
    m = n_eq + n_ineq
    params = _layout_params([('x_init', (nx,), x_init)])
    n_theta = nx + 1

    # --- P (upper triangle, CSC) and A (CSC): constants of the family
    P = sp.block_diag([sp.csc_matrix(2 * Q)] * N + [sp.csc_matrix(2 * QN)] + [sp.csc_matrix(2 * R)] * N, format='csc')
    Pu = sp.triu(P, format='csc'); Pu.eliminate_zeros(); Pu.sort_indices()
    Ax_blocks = sp.eye(nX, format='csc') - sp.kron(sp.eye(N + 1, k=-1), sp.csc_matrix(Ad), format='csc')
    Au_blocks = -sp.kron(sp.vstack([sp.csc_matrix((1, N)), sp.eye(N)]), sp.csc_matrix(Bd), format='csc')
    A = sp.vstack([sp.hstack([Ax_blocks, Au_blocks]),
                   sp.hstack([sp.csc_matrix((nU, nX)), sp.eye(nU)])], format='csc')
    A.eliminate_zeros(); A.sort_indices()

    maps = {}
    mb = _MapBuilder(Pu.nnz, n_theta)
    for k, v in enumerate(Pu.data):
        mb.const(k, v)
    maps['P'] = mb.csr()
    mb = _MapBuilder(A.nnz, n_theta)
    for k, v in enumerate(A.data):
        mb.const(k, v)
    maps['A'] = mb.csr()
    maps['q'] = sp.csr_matrix((n, n_theta))
    mb = _MapBuilder(1, n_theta); mb.const(0, d_const); maps['d'] = mb.csr()
    ml, mu = _MapBuilder(m, n_theta), _MapBuilder(m, n_theta)
    for i in range(nx):                      # x_0 = x_init
        ml.add(i, i, 1.0); mu.add(i, i, 1.0)
    for i in range(n_ineq):                  # input box
        ml.const(n_eq + i, -umax); mu.const(n_eq + i, umax)
    maps['l'], maps['u'] = ml.csr(), mu.csr()

    variables = [UserVar('U', (nu, N), nX + np.arange(nU)),
                 UserVar('X', (nx, N + 1), np.arange(nX))]
    duals = [UserDual('d0', 'y', (nx, N), nx + np.arange(N * nx)),
             UserDual('d1', 'y', (nu, N), n_eq + np.arange(nU)),
             UserDual('d2', 'y', (nx,), np.arange(nx))]
    return CanonFamily(name or f'mpc_{nx}_{nu}_{N}', 'quadratic', n, n_eq, n_ineq, params, maps,
                       {'P': _csc_pattern(Pu), 'A': _csc_pattern(A)}, variables, duals)


def mpc_ltv(nx=12, nu=4, N=10, Ad=None, Bd=None, qdiag=None, rdiag=None, umax=1.0, x_init=None,
            name=None) -> CanonFamily:
    """The MPC QP of ``mpc`` with the reference family's MATRIX parameters (tests/test_E2E_QP.py:54-61 declares the
    dynamics and the cost factors as cvxpy Parameters): ``A`` (nx x nx), ``B`` (nx x nu), the diagonal stage costs
    ``qdiag`` (nx) and ``rdiag`` (nu), and ``x_init``.  ``A``/``B`` enter the canonical constraint matrix (every one of
    their entries appears in each of the N stages), ``qdiag``/``rdiag`` the canonical cost matrix P = 2*blkdiag(diag(qdiag)
    x (N+1), diag(rdiag) x N) -- so batching them exercises the osqp_update_data_mat branch of the generated solve
    (cvxpygen/solvers/osqp.py:20-33): per-instance re-scaling, KKT assembly and numeric LDL' (SURVEY row f2).
    The sparsity PATTERN is that of dense A and B: structural entries stay in the pattern even when a value is zero."""
    dA, dB = default_mpc_dynamics(nx, nu)
    Ad = dA if Ad is None else np.asarray(Ad, float)
    Bd = dB if Bd is None else np.asarray(Bd, float)
    qdiag = np.ones(nx) if qdiag is None else np.asarray(qdiag, float)
    rdiag = 0.1 * np.ones(nu) if rdiag is None else np.asarray(rdiag, float)
    x_init = np.zeros(nx) if x_init is None else np.asarray(x_init, float)
    nX, nU = (N + 1) * nx, N * nu
    n, n_eq, n_ineq = nX + nU, nX, nU
    m = n_eq + n_ineq
    params = _layout_params([('A', (nx, nx), Ad.flatten(order='F')), ('B', (nx, nu), Bd.flatten(order='F')),
                             ('qdiag', (nx,), qdiag), ('rdiag', (nu,), rdiag), ('x_init', (nx,), x_init)])
    col = {p.name: p.col for p in params}
    n_theta = params[-1].col + params[-1].size + 1
    # structural patterns: P diagonal; A = [[I - shift(A), -shift(B)], [0, I]] with dense A and B blocks
    Pu = sp.identity(n, format='csc')
    rows, cols, kind = [], [], []          # kind: ('one',) | ('A', i, j) | ('B', i, j)
    for i in range(nX):
        rows.append(i); cols.append(i); kind.append(('one',))
    for k in range(N):
        for i in range(nx):
            for j in range(nx):
                rows.append((k + 1) * nx + i); cols.append(k * nx + j); kind.append(('A', i, j))
            for j in range(nu):
                rows.append((k + 1) * nx + i); cols.append(nX + k * nu + j); kind.append(('B', i, j))
    for i in range(nU):
        rows.append(n_eq + i); cols.append(nX + i); kind.append(('one',))
    tag = sp.csc_matrix((np.arange(1, len(rows) + 1, dtype=float), (rows, cols)), shape=(m, n))
    tag.sort_indices()
    order = tag.data.astype(int) - 1                  # CSC position -> triplet number
    mb = _MapBuilder(len(order), n_theta)
    for pos, t in enumerate(order):
        kd = kind[t]
        if kd[0] == 'one':
            mb.const(pos, 1.0)
        elif kd[0] == 'A':
            mb.add(pos, col['A'] + kd[1] + nx * kd[2], -1.0)
        else:
            mb.add(pos, col['B'] + kd[1] + nx * kd[2], -1.0)
    maps = {'A': mb.csr()}
    mp = _MapBuilder(n, n_theta)
    for i in range(nX):
        mp.add(i, col['qdiag'] + (i % nx), 2.0)
    for i in range(nU):
        mp.add(nX + i, col['rdiag'] + (i % nu), 2.0)
    maps['P'] = mp.csr()
    maps['q'] = sp.csr_matrix((n, n_theta))
    md = _MapBuilder(1, n_theta); md.const(0, 0.0); maps['d'] = md.csr()
    ml, mu = _MapBuilder(m, n_theta), _MapBuilder(m, n_theta)
    for i in range(nx):
        ml.add(i, col['x_init'] + i, 1.0); mu.add(i, col['x_init'] + i, 1.0)
    for i in range(n_ineq):
        ml.const(n_eq + i, -umax); mu.const(n_eq + i, umax)
    maps['l'], maps['u'] = ml.csr(), mu.csr()
    variables = [UserVar('U', (nu, N), nX + np.arange(nU)), UserVar('X', (nx, N + 1), np.arange(nX))]
    duals = [UserDual('d0', 'y', (nx, N), nx + np.arange(N * nx)), UserDual('d1', 'y', (nu, N), n_eq + np.arange(nU)),
             UserDual('d2', 'y', (nx,), np.arange(nx))]
    Apat = (tag.indices.astype(np.int32), tag.indptr.astype(np.int32), (m, n))
    return CanonFamily(name or f'mpc_ltv_{nx}_{nu}_{N}', 'quadratic', n, n_eq, n_ineq, params, maps,
                       {'P': _csc_pattern(Pu), 'A': Apat}, variables, duals)


def mpc_reference(H=10, name=None) -> CanonFamily:
    """The reference's own test MPC (tests/test_E2E_QP.py:44-73, data :127-146) with ALL its parameters: n = 6 states,
    m = 3 inputs, horizon H;  ``Psqrt``, ``Qsqrt``, ``Rsqrt`` diagonal parameters (stored: the diagonal), ``A`` and ``B``
    SPARSE parameters (stored: the nonzeros in column-major order, reference README.md:140-143), ``x_init``.

        min  |Psqrt X[:,H-1]|^2 + |Qsqrt X[:,:H]|^2 + |Rsqrt U|^2 + 1
        s.t. X[:,1:] = A X[:,:H] + B U,  |U| <= 1,  X[:,0] = x_init

    The cost factors multiply variables inside sum_squares, so -- as in cvxpy's DPP canonicalisation -- they enter through
    auxiliary variables: canonical x = [X(:) ; U(:) ; TQ = Qsqrt X[:,:H] ; TP = Psqrt X[:,H-1] ; TR = Rsqrt U], P = 2 I on
    the auxiliaries (constant), every matrix parameter in the constraint matrix A; objective offset d = 1.
    rows: [x_0 = x_init (d2) ; dynamics (d0) ; TQ, TP, TR definitions ; -1 <= U <= 1 (d1)]."""
    n, m = 6, 3
    nzA = sorted([(i, i) for i in range(n)] + [(i, 3 + i) for i in range(n // 2)], key=lambda rc: (rc[1], rc[0]))
    nzB = sorted([(3 + i, i) for i in range(n // 2)], key=lambda rc: (rc[1], rc[0]))
    td = 0.1
    A0 = np.eye(6); A0[:3, 3:] += td * np.eye(3)
    B0 = np.zeros((6, 3)); B0[3:, :] = td * np.eye(3)
    params = _layout_params([('Psqrt', (n, n), np.ones(n)), ('Qsqrt', (n, n), np.ones(n)),
                             ('Rsqrt', (m, m), np.sqrt(0.1) * np.ones(m)),
                             ('A', (n, n), [A0[r, c] for r, c in nzA]), ('B', (n, m), [B0[r, c] for r, c in nzB]),
                             ('x_init', (n,), np.zeros(n))])
    col = {p.name: p.col for p in params}
    n_theta = params[-1].col + params[-1].size + 1
    nX, nU = n * (H + 1), m * H
    oX, oU, oTQ, oTP, oTR = 0, nX, nX + nU, nX + nU + n * H, nX + nU + n * H + n
    nv = oTR + m * H
    r0, rD, rTQ, rTP, rTR, rB = 0, n, n + n * H, n + 2 * n * H, 2 * n + 2 * n * H, 2 * n + 2 * n * H + m * H
    n_eq, n_ineq = rB, m * H
    mt = n_eq + n_ineq
    ent = []                                         # (row, col, ('c', value) | ('p', theta column, coefficient))
    for i in range(n):
        ent.append((r0 + i, oX + i, ('c', 1.0)))
    for k in range(H):
        for i in range(n):
            ent.append((rD + k * n + i, oX + (k + 1) * n + i, ('c', 1.0)))
            ent.append((rTQ + k * n + i, oTQ + k * n + i, ('c', 1.0)))
            ent.append((rTQ + k * n + i, oX + k * n + i, ('p', col['Qsqrt'] + i, -1.0)))
        for e, (r, c) in enumerate(nzA):
            ent.append((rD + k * n + r, oX + k * n + c, ('p', col['A'] + e, -1.0)))
        for e, (r, c) in enumerate(nzB):
            ent.append((rD + k * n + r, oU + k * m + c, ('p', col['B'] + e, -1.0)))
        for j in range(m):
            ent.append((rTR + k * m + j, oTR + k * m + j, ('c', 1.0)))
            ent.append((rTR + k * m + j, oU + k * m + j, ('p', col['Rsqrt'] + j, -1.0)))
            ent.append((rB + k * m + j, oU + k * m + j, ('c', 1.0)))
    for i in range(n):
        ent.append((rTP + i, oTP + i, ('c', 1.0)))
        ent.append((rTP + i, oX + (H - 1) * n + i, ('p', col['Psqrt'] + i, -1.0)))
    ent.sort(key=lambda e: (e[1], e[0]))
    Ar = np.array([e[0] for e in ent]); Ac = np.array([e[1] for e in ent])
    indptr = np.zeros(nv + 1, dtype=np.int64)
    np.add.at(indptr, Ac + 1, 1); indptr = np.cumsum(indptr).astype(np.int32)
    mbA = _MapBuilder(len(ent), n_theta)
    for k, e in enumerate(ent):
        if e[2][0] == 'c':
            mbA.const(k, e[2][1])
        else:
            mbA.add(k, e[2][1], e[2][2])
    naux = nv - oTQ
    Pu = sp.csc_matrix((2.0 * np.ones(naux), (oTQ + np.arange(naux), oTQ + np.arange(naux))), shape=(nv, nv))
    mbP = _MapBuilder(naux, n_theta)
    for k in range(naux):
        mbP.const(k, 2.0)
    maps = {'A': mbA.csr(), 'P': mbP.csr(), 'q': sp.csr_matrix((nv, n_theta))}
    md = _MapBuilder(1, n_theta); md.const(0, 1.0); maps['d'] = md.csr()
    ml, mu = _MapBuilder(mt, n_theta), _MapBuilder(mt, n_theta)
    for i in range(n):
        ml.add(r0 + i, col['x_init'] + i, 1.0); mu.add(r0 + i, col['x_init'] + i, 1.0)
    for i in range(n_ineq):
        ml.const(rB + i, -1.0); mu.const(rB + i, 1.0)
    maps['l'], maps['u'] = ml.csr(), mu.csr()
    variables = [UserVar('U', (m, H), oU + np.arange(nU)), UserVar('X', (n, H + 1), oX + np.arange(nX))]
    duals = [UserDual('d0', 'y', (n, H), rD + np.arange(n * H)), UserDual('d1', 'y', (m, H), rB + np.arange(m * H)),
             UserDual('d2', 'y', (n,), r0 + np.arange(n))]
    return CanonFamily(name or f'mpc_ref_6_3_{H}', 'quadratic', nv, n_eq, n_ineq, params, maps,
                       {'P': _csc_pattern(Pu), 'A': (Ar.astype(np.int32), indptr, (mt, nv))}, variables, duals)


def mpc_reference_batch(fam: CanonFamily, B: int, seed: int = 0, spread: float = 0.05):
    """Per-instance values of all six parameters: the reference's data (tests/test_E2E_QP.py:127-146: x_init = -2 + 4 rand)
    with the stored entries of A, B and the cost factors perturbed entry-wise (relative N(0, spread^2))."""
    rng = np.random.default_rng(seed)
    out = {}
    for nm in ('Psqrt', 'Qsqrt', 'Rsqrt', 'A', 'B'):
        d = fam.param(nm).default
        out[nm] = d[None, :] * (1.0 + spread * rng.standard_normal((B, d.size)))
    out['x_init'] = -2.0 + 4.0 * rng.random((B, 6))
    return out


def mpc_ltv_batch(fam: CanonFamily, B: int, seed: int = 3, spread: float = 0.05):
    """Synthetic per-instance parameters for ``mpc_ltv`` (bench.py --workload mpc_ltv and the tests): dynamics perturbed
    entry-wise around the family's defaults (N(0, spread^2)), diagonal stage costs U[0.5, 2] / U[0.05, 0.5], x_init U[-1, 1]."""
    rng = np.random.default_rng(seed)
    A0, B0 = fam.param('A').default, fam.param('B').default
    return {'A': A0[None, :] + spread * rng.standard_normal((B, A0.size)),
            'B': B0[None, :] + spread * rng.standard_normal((B, B0.size)),
            'qdiag': rng.uniform(0.5, 2.0, (B, fam.param('qdiag').size)),
            'rdiag': rng.uniform(0.05, 0.5, (B, fam.param('rdiag').size)),
            'x_init': rng.uniform(-1, 1, (B, fam.param('x_init').size))}


def nonneg_ls(m=3, n=2, A_pattern=None, A_data=None, b=None, seed=1, name=None) -> CanonFamily:
    """README example  min ||A x - b||^2  s.t. x >= 0  (reference: examples/main.py:16-26).

    canonical x = [x(n) ; t(m)] with t = A x - b (the epigraph-free part of what cvxpy's
    sum_squares canonicalisation introduces), P = 2 I on t.
    rows:  [0, m)      A x - t = b          (no user dual)
           [m, m+n)    -x <= 0              (user constraint d0, dual >= 0)
    User parameters: ``A`` (sparse, stored entries in column order -- reference README.md:140-143)
    enters the canonical matrix A; ``b`` enters l and u.
    """
    rs = np.random.RandomState(seed)
    if A_pattern is None:
        A_pattern = ((0, 0, 1), (0, 1, 1)) if (m, n) == (3, 2) else tuple(np.nonzero(np.ones((m, n))))
    rows, cols = np.asarray(A_pattern[0]), np.asarray(A_pattern[1])
    order = np.lexsort((rows, cols))            # column-major order of the stored entries
    rows, cols = rows[order], cols[order]
    nnzA = rows.size
    if A_data is None:
        A_data = rs.randn(nnzA)
    if b is None:
        b = rs.randn(m)
    params = _layout_params([('A', (m, n), A_data), ('b', (m,), b)])
    colA, colb = params[0].col, params[1].col
    n_theta = nnzA + m + 1
    nv, n_eq, n_ineq = n + m, m, n
    mtot = n_eq + n_ineq

    Pu = sp.csc_matrix((2 * np.ones(m), (n + np.arange(m), n + np.arange(m))), shape=(nv, nv))
    # canonical A pattern: [A_user, -I ; -I, 0]; tag each stored entry with its source
    ent = [(r, c, ('A', k)) for k, (r, c) in enumerate(zip(rows, cols))]
    ent += [(i, n + i, ('c', -1.0)) for i in range(m)]
    ent += [(m + j, j, ('c', -1.0)) for j in range(n)]
    ent.sort(key=lambda e: (e[1], e[0]))
    Ar = np.array([e[0] for e in ent]); Ac = np.array([e[1] for e in ent])
    indptr = np.zeros(nv + 1, dtype=np.int32)
    np.add.at(indptr, Ac + 1, 1); indptr = np.cumsum(indptr).astype(np.int32)
    mbA = _MapBuilder(len(ent), n_theta)
    for k, e in enumerate(ent):
        if e[2][0] == 'A':
            mbA.add(k, colA + e[2][1], 1.0)
        else:
            mbA.const(k, e[2][1])
    maps = {'A': mbA.csr()}
    mb = _MapBuilder(Pu.nnz, n_theta)
    for k in range(Pu.nnz):
        mb.const(k, 2.0)
    maps['P'] = mb.csr()
    maps['q'] = sp.csr_matrix((nv, n_theta))
    maps['d'] = sp.csr_matrix((1, n_theta))
    ml, mu = _MapBuilder(mtot, n_theta), _MapBuilder(mtot, n_theta)
    for i in range(m):
        ml.add(i, colb + i, 1.0); mu.add(i, colb + i, 1.0)
    for j in range(n):
        ml.const(m + j, -INF)
    maps['l'], maps['u'] = ml.csr(), mu.csr()
    variables = [UserVar('x', (n,), np.arange(n))]
    duals = [UserDual('d0', 'y', (n,), m + np.arange(n))]
    return CanonFamily(name or f'nonneg_LS_{m}_{n}', 'quadratic', nv, n_eq, n_ineq, params, maps,
                       {'P': _csc_pattern(Pu), 'A': (Ar.astype(np.int32), indptr, (mtot, nv))},
                       variables, duals)


def random_qp(n=20, m_eq=5, m_ineq=15, density=0.3, seed=0, name=None) -> CanonFamily:
    """Generic sparse QP family with unstructured sparsity -- exercises the generic level scheduler.
    User parameters: ``q`` (linear cost), ``b`` (equality rhs), ``h`` (inequality upper bounds)."""
    rs = np.random.RandomState(seed)
    Mh = sp.random(n, n, density=density, random_state=rs, data_rvs=rs.randn).toarray()
    Pfull = Mh @ Mh.T * 0.5 + 0.1 * np.eye(n)
    Pfull[np.abs(Pfull) < 0.05] = 0.0
    Pfull = 0.5 * (Pfull + Pfull.T) + np.eye(n) * (np.abs(Pfull).sum(1).max())  # diagonally dominant => PSD
    Pu = sp.triu(sp.csc_matrix(Pfull), format='csc'); Pu.sort_indices()
    Aeq = sp.random(m_eq, n, density=density, random_state=rs, data_rvs=rs.randn).toarray()
    for i in range(m_eq):
        if not Aeq[i].any():
            Aeq[i, rs.randint(n)] = 1.0
    Ain = sp.random(m_ineq, n, density=density, random_state=rs, data_rvs=rs.randn).toarray()
    for i in range(m_ineq):
        if not Ain[i].any():
            Ain[i, rs.randint(n)] = 1.0
    A = sp.csc_matrix(np.vstack([Aeq, Ain])); A.sort_indices()
    x0 = rs.randn(n)
    params = _layout_params([('q', (n,), rs.randn(n)), ('b', (m_eq,), Aeq @ x0),
                             ('h', (m_ineq,), Ain @ x0 + rs.rand(m_ineq))])
    n_theta = n + m_eq + m_ineq + 1
    m = m_eq + m_ineq
    maps = {}
    mb = _MapBuilder(Pu.nnz, n_theta)
    for k, v in enumerate(Pu.data):
        mb.const(k, v)
    maps['P'] = mb.csr()
    mb = _MapBuilder(A.nnz, n_theta)
    for k, v in enumerate(A.data):
        mb.const(k, v)
    maps['A'] = mb.csr()
    mb = _MapBuilder(n, n_theta)
    for i in range(n):
        mb.add(i, params[0].col + i, 1.0)
    maps['q'] = mb.csr()
    maps['d'] = sp.csr_matrix((1, n_theta))
    ml, mu = _MapBuilder(m, n_theta), _MapBuilder(m, n_theta)
    for i in range(m_eq):
        ml.add(i, params[1].col + i, 1.0); mu.add(i, params[1].col + i, 1.0)
    for i in range(m_ineq):
        ml.const(m_eq + i, -INF); mu.add(m_eq + i, params[2].col + i, 1.0)
    maps['l'], maps['u'] = ml.csr(), mu.csr()
    variables = [UserVar('x', (n,), np.arange(n))]
    duals = [UserDual('d0', 'y', (m_eq,), np.arange(m_eq)), UserDual('d1', 'y', (m_ineq,), m_eq + np.arange(m_ineq))]
    return CanonFamily(name or f'random_qp_{n}_{m_eq}_{m_ineq}', 'quadratic', n, m_eq, m_ineq, params, maps,
                       {'P': _csc_pattern(Pu), 'A': _csc_pattern(A)}, variables, duals)


def portfolio_socp(n=100, m=10, seed=1, k_tc=0.01, k_sh=0.05, Lmax=1.6, name=None, matrix_params=False) -> CanonFamily:
    """Portfolio optimisation as an SOCP in ECOS form (reference problem:
    examples/portfolio.ipynb / tests/test_E2E_QP.py:76-110,148-162):

        maximise a'w - ||Sig_f^1/2 f||^2 - ||d o w||^2 - k_tc'|dw| + k_sh' min(0, w)
        s.t.     f = F'w,  1'w = 1,  ||w||_1 <= L,  dw = w - w_prev

    canonical x = [w(n), dw(n), f(m), t_tc(n), t_sh(n), t_l1(n), r1, r2];   min c'x  s.t.  A x = b,  h - G x in K
      equalities (p = m + 1 + n):  f - F'w = 0 ;  1'w = 1 ;  dw - w = -w_prev
      LP cone (l = 6n + 1):        t_tc -/+ dw >= 0 ; t_sh + w >= 0, t_sh >= 0 ; t_l1 -/+ w >= 0 ; L - 1't_l1 >= 0
      SOC(m+2): (1 + r1, 1 - r1, 2 Sig_f^1/2 f) ;  SOC(n+2): (1 + r2, 1 - r2, 2 d o w)      (||v||^2 <= r)
    User parameters: ``a`` (enters c), ``w_prev`` (enters b); F, Sig_f_sqrt, d_sqrt, k_tc, k_sh, L are constants here.
    ``matrix_params=True`` declares the factor loadings ``F`` (n x m, column-major like cvxpy) and ``d_sqrt`` (n) as user
    parameters as the reference's example does (examples/portfolio.ipynb): F enters A (every entry is structural, zero or
    not), d_sqrt enters G -- the per-instance matrix path of IPM-CUDA (ECOS_updateData with new G / A values).
    Maximisation: obj_val is negated at retrieval (cvxpygen/utils.py:980)."""
    rs = np.random.RandomState(seed)
    alpha = rs.randn(n)
    F = np.round(rs.randn(n, m))
    sig = rs.rand(m)
    d = rs.rand(n)
    nv = 5 * n + m + 2
    iw, idw, if_, itc, ish, il1, ir1, ir2 = 0, n, 2 * n, 2 * n + m, 3 * n + m, 4 * n + m, 5 * n + m, 5 * n + m + 1
    p = m + 1 + n
    specs = [('a', (n,), alpha), ('w_prev', (n,), np.zeros(n))]
    if matrix_params:
        specs += [('F', (n, m), F.flatten(order='F')), ('d_sqrt', (n,), d)]
    params = _layout_params(specs)
    col_a, col_wp = params[0].col, params[1].col
    col_F, col_d = (params[2].col, params[3].col) if matrix_params else (-1, -1)
    n_theta = params[-1].col + params[-1].size + 1
    # ---- A x = b
    Ar, Ac, Av = [], [], []
    Fent = {}                                      # (row, col) of A -> (i, j) of F
    for j in range(m):
        Ar.append(j); Ac.append(if_ + j); Av.append(1.0)
        for i in range(n):
            if F[i, j] != 0 or matrix_params:
                Ar.append(j); Ac.append(iw + i); Av.append(-F[i, j] if F[i, j] != 0 else 1.0)     # structural entry: any nonzero
                Fent[(j, iw + i)] = (i, j)
    for i in range(n):
        Ar.append(m); Ac.append(iw + i); Av.append(1.0)
    for i in range(n):
        Ar += [m + 1 + i, m + 1 + i]; Ac += [idw + i, iw + i]; Av += [1.0, -1.0]
    A = sp.csc_matrix((Av, (Ar, Ac)), shape=(p, nv)); A.sort_indices()
    # ---- h - G x in K
    Gr, Gc, Gv, hv = [], [], [], []
    row = 0

    def ge(terms, const=0.0):      # sum coef*x + const >= 0   ->  G row = -coef, h = const
        nonlocal row
        for cidx, coef in terms:
            Gr.append(row); Gc.append(cidx); Gv.append(-coef)
        hv.append(const); row += 1
    for i in range(n): ge([(itc + i, 1.0), (idw + i, -1.0)])
    for i in range(n): ge([(itc + i, 1.0), (idw + i, 1.0)])
    for i in range(n): ge([(ish + i, 1.0), (iw + i, 1.0)])
    for i in range(n): ge([(ish + i, 1.0)])
    for i in range(n): ge([(il1 + i, 1.0), (iw + i, -1.0)])
    for i in range(n): ge([(il1 + i, 1.0), (iw + i, 1.0)])
    ge([(il1 + i, -1.0) for i in range(n)], Lmax)
    n_lp = row
    ge([(ir1, 1.0)], 1.0); ge([(ir1, -1.0)], 1.0)
    for j in range(m): ge([(if_ + j, 2.0 * sig[j])])
    ge([(ir2, 1.0)], 1.0); ge([(ir2, -1.0)], 1.0)
    d_rows = {}
    for i in range(n):
        d_rows[(row, iw + i)] = i
        ge([(iw + i, 2.0 * d[i])])
    mc = row
    G = sp.csc_matrix((Gv, (Gr, Gc)), shape=(mc, nv)); G.sort_indices()
    maps = {}
    mb = _MapBuilder(nv, n_theta)
    for i in range(n):
        mb.add(iw + i, col_a + i, -1.0); mb.const(itc + i, k_tc); mb.const(ish + i, k_sh)
    mb.const(ir1, 1.0); mb.const(ir2, 1.0)
    maps['c'] = mb.csr()
    maps['d'] = sp.csr_matrix((1, n_theta))
    mbA = _MapBuilder(A.nnz, n_theta)
    Acols = np.repeat(np.arange(nv), np.diff(A.indptr))
    for k, v in enumerate(A.data):
        key = (int(A.indices[k]), int(Acols[k]))
        if matrix_params and key in Fent:
            i, j = Fent[key]
            mbA.add(k, col_F + j * n + i, -1.0)            # A[j, w_i] = -F[i, j]
        else:
            mbA.const(k, v)
    maps['A'] = mbA.csr()
    mbG = _MapBuilder(G.nnz, n_theta)
    Gcols = np.repeat(np.arange(nv), np.diff(G.indptr))
    for k, v in enumerate(G.data):
        key = (int(G.indices[k]), int(Gcols[k]))
        if matrix_params and key in d_rows:
            mbG.add(k, col_d + d_rows[key], -2.0)           # G row of the second cone: -(2 d_i) w_i
        else:
            mbG.const(k, v)
    maps['G'] = mbG.csr()
    mbb = _MapBuilder(p, n_theta); mbb.const(m, 1.0)
    for i in range(n): mbb.add(m + 1 + i, col_wp + i, -1.0)
    maps['b'] = mbb.csr()
    mbh = _MapBuilder(mc, n_theta)
    for k, v in enumerate(hv):
        if v != 0.0: mbh.const(k, v)
    maps['h'] = mbh.csr()
    variables = [UserVar('w', (n,), iw + np.arange(n)), UserVar('delta_w', (n,), idw + np.arange(n)),
                 UserVar('f', (m,), if_ + np.arange(m))]
    duals = [UserDual('d0', 'y', (m,), np.arange(m)), UserDual('d1', 'y', (1,), np.array([m])),
             UserDual('d2', 'z', (1,), np.array([n_lp - 1])), UserDual('d3', 'y', (n,), m + 1 + np.arange(n))]
    return CanonFamily(name or (f'portfolio_socp_mat_{n}_{m}' if matrix_params else f'portfolio_socp_{n}_{m}'), 'conic', nv, p, mc, params, maps,
                       {'A': _csc_pattern(A), 'G': _csc_pattern(G)}, variables, duals, is_maximization=True,
                       cone_dims={'l': n_lp, 'q': [m + 2, n + 2]})


def random_socp(n=30, p=8, l=20, q=(3, 5, 4), density=0.25, seed=5, name=None) -> CanonFamily:
    """Generic conic family in ECOS form with every vector a user parameter:

        minimise c'x   s.t.  A x = b,   h - G x in  R+^l x SOC(q_1) x ... x SOC(q_k)

    (the canonical form the reference hands to ECOS for any DPP SOCP, cvxpygen/solvers/ecos.py:45-58; the reference's own
    conic tests are small SOCPs / LPs of this shape, tests/test_E2E_SOCP.py:15-64, tests/test_E2E_LP.py).  A, G are
    random sparse constants; the defaults of c, b, h are built from a strictly feasible primal-dual pair, so the nominal
    instance is solvable, while shifted b / h / c give primal-infeasible and unbounded instances for the exit-flag tests.
    q = () gives a pure LP (no second-order cone), p = 0 a problem without equalities."""
    rs = np.random.RandomState(seed)
    q = [int(d) for d in q]
    m = l + sum(q)

    def sprand(r, c_):
        M = sp.random(r, c_, density=density, random_state=rs, data_rvs=rs.randn).tolil()
        for i in range(r):                     # no empty rows: every constraint involves a variable
            if M.rows[i] == []:
                M[i, rs.randint(c_)] = rs.randn()
        return sp.csc_matrix(M)
    A = sprand(p, n) if p else sp.csc_matrix((0, n))
    G = sprand(m, n)
    for j in range(n):                          # no empty columns either: every variable is constrained
        if G[:, j].nnz == 0:
            G = G.tolil(); G[rs.randint(m), j] = rs.randn(); G = sp.csc_matrix(G)
    # structure for the exit-flag tests: LP rows 0 and 1 are  g'x <= h_0  and  -g'x <= h_1  (infeasible when h_0 + h_1 < 0);
    # the last variable only appears in LP row 2 as  -x_last <= h_2  (unbounded when its cost is negative)
    assert l >= 3
    G = G.tolil(); A = A.tolil()
    G[1, :] = -G[0, :].toarray()
    G[:, n - 1] = 0.0; G[2, :] = 0.0; G[2, n - 1] = -1.0
    if p:
        A[:, n - 1] = 0.0
        for i in range(p):
            if A.rows[i] == []:
                A[i, rs.randint(n - 1)] = rs.randn()
    for i in range(m):
        if G.rows[i] == []:
            G[i, rs.randint(n - 1)] = rs.randn()
    A = sp.csc_matrix(A); G = sp.csc_matrix(G); A.eliminate_zeros(); G.eliminate_zeros()
    A.sort_indices(); G.sort_indices()
    x0 = rs.randn(n)
    s0 = np.r_[0.5 + rs.rand(l)]
    z0 = np.r_[0.5 + rs.rand(l)]
    for d in q:
        v = rs.randn(d - 1); s0 = np.r_[s0, np.linalg.norm(v) + 0.5 + rs.rand(), v]
        v = rs.randn(d - 1); z0 = np.r_[z0, np.linalg.norm(v) + 0.5 + rs.rand(), v]
    h0 = G @ x0 + s0
    b0 = A @ x0 if p else np.zeros(0)
    c0 = -(A.T @ rs.randn(p) if p else 0.0) - G.T @ z0
    specs = [('c', (n,), c0)] + ([('b', (p,), b0)] if p else []) + [('h', (m,), h0)]
    params = _layout_params(specs)
    n_theta = sum(pp.size for pp in params) + 1
    col = {pp.name: pp.col for pp in params}
    maps = {}
    for pid, size in (('c', n), ('b', p), ('h', m)):
        mb = _MapBuilder(size, n_theta)
        if size and pid in col:
            for i in range(size):
                mb.add(i, col[pid] + i, 1.0)
        maps[pid] = mb.csr()
    maps['d'] = sp.csr_matrix((1, n_theta))
    for pid, M in (('A', A), ('G', G)):
        mb = _MapBuilder(M.nnz, n_theta)
        for k, v in enumerate(M.data):
            mb.const(k, v)
        maps[pid] = mb.csr()
    variables = [UserVar('x', (n,), np.arange(n))]
    duals = ([UserDual('d0', 'y', (p,), np.arange(p))] if p else []) + [UserDual('d1' if p else 'd0', 'z', (m,), np.arange(m))]
    tag = 'x'.join(str(d) for d in q) if q else 'lp'
    return CanonFamily(name or f'random_socp_{n}_{p}_{l}_{tag}', 'conic', n, p, m, params, maps,
                       {'A': _csc_pattern(A), 'G': _csc_pattern(G)}, variables, duals, is_maximization=False,
                       cone_dims={'l': l, 'q': q})


def actuator(name=None) -> CanonFamily:
    """The reference's actuator-allocation test problem (tests/test_E2E_QP.py:16-41, data :116-125), chosen there "for
    degenerate vectors and matrices": one actuator (n = 1), three objectives (m = 3),

        minimise |A u - w|^2 + lamb_sm |delta_u|^2 + kappa'|u|   s.t.  u_min <= u <= u_max,  delta_u = u - u_prev

    canonical x = [u ; delta_u ; r = A u - w (3) ; t >= |u|];  P = diag(0, 2 lamb_sm, 2, 2, 2, 0) -- the scalar parameter
    lamb_sm enters P --, q = kappa on t;  rows: r - A u = -w | delta_u - u = -u_prev (d2) | u_min <= u <= u_max (d0 / d1 share
    the two-sided row) | +-u - t <= 0.  Scalar parameters have shape ()."""
    params = _layout_params([('A', (3, 1), np.ones(3)), ('w', (3,), np.array([2.0, 3.0, 5.0])), ('lamb_sm', (), 0.5488135039273248),
                             ('kappa', (1,), 0.1), ('u_prev', (1,), 0.0), ('u_min', (1,), -1.0), ('u_max', (1,), 1.0)])
    col = {p.name: p.col for p in params}
    n_theta = params[-1].col + params[-1].size + 1
    iu, idu, ir, it = 0, 1, 2, 5
    nv = 6
    ent = [(0 + i, ir + i, ('c', 1.0)) for i in range(3)] + [(0 + i, iu, ('p', col['A'] + i, -1.0)) for i in range(3)]
    ent += [(3, idu, ('c', 1.0)), (3, iu, ('c', -1.0)), (4, iu, ('c', 1.0)),
            (5, iu, ('c', 1.0)), (5, it, ('c', -1.0)), (6, iu, ('c', -1.0)), (6, it, ('c', -1.0))]
    n_eq, mt = 4, 7
    ent.sort(key=lambda e: (e[1], e[0]))
    Ar = np.array([e[0] for e in ent]); Ac = np.array([e[1] for e in ent])
    indptr = np.zeros(nv + 1, dtype=np.int64)
    np.add.at(indptr, Ac + 1, 1); indptr = np.cumsum(indptr).astype(np.int32)
    mbA = _MapBuilder(len(ent), n_theta)
    for k, e in enumerate(ent):
        if e[2][0] == 'c':
            mbA.const(k, e[2][1])
        else:
            mbA.add(k, e[2][1], e[2][2])
    Pu = sp.csc_matrix((np.ones(4), ([idu, ir, ir + 1, ir + 2], [idu, ir, ir + 1, ir + 2])), shape=(nv, nv))
    mbP = _MapBuilder(4, n_theta)
    mbP.add(0, col['lamb_sm'], 2.0)
    for k in range(1, 4):
        mbP.const(k, 2.0)
    mq = _MapBuilder(nv, n_theta); mq.add(it, col['kappa'], 1.0)
    ml, mu = _MapBuilder(mt, n_theta), _MapBuilder(mt, n_theta)
    for i in range(3):
        ml.add(i, col['w'] + i, -1.0); mu.add(i, col['w'] + i, -1.0)
    ml.add(3, col['u_prev'], -1.0); mu.add(3, col['u_prev'], -1.0)
    ml.add(4, col['u_min'], 1.0); mu.add(4, col['u_max'], 1.0)
    ml.const(5, -INF); ml.const(6, -INF)
    maps = {'A': mbA.csr(), 'P': mbP.csr(), 'q': mq.csr(), 'd': sp.csr_matrix((1, n_theta)), 'l': ml.csr(), 'u': mu.csr()}
    variables = [UserVar('u', (1,), np.array([iu])), UserVar('delta_u', (1, 1), np.array([idu]))]
    duals = [UserDual('d0', 'y', (1,), np.array([4])), UserDual('d1', 'y', (1, 1), np.array([3]))]
    return CanonFamily(name or 'actuator_1_3', 'quadratic', nv, n_eq, mt - n_eq, params, maps,
                       {'P': _csc_pattern(Pu), 'A': (Ar.astype(np.int32), indptr, (mt, nv))}, variables, duals)


def actuator_batch(fam: CanonFamily, B: int, seed: int = 0):
    """Per-instance values of ALL seven parameters (lamb_sm enters P, A enters the constraint matrix)."""
    rng = np.random.default_rng(seed)
    lo = -1.0 - 0.5 * rng.random((B, 1))
    return {'A': 1.0 + 0.3 * rng.standard_normal((B, 3)), 'w': np.array([2.0, 3.0, 5.0]) + rng.standard_normal((B, 3)),
            'lamb_sm': rng.random((B, 1)), 'kappa': 0.1 + 0.1 * rng.random((B, 1)), 'u_prev': 0.5 * rng.standard_normal((B, 1)),
            'u_min': lo, 'u_max': lo + 0.5 + 2.0 * rng.random((B, 1))}


def osqp_update_matrices_kat(name=None):
    """OSQP's own known-answer test for matrix updates (osqp_sources/tests/update_matrices/generate_problem.py: n = 5, m = 8,
    numpy Generator(PCG64(2)) -- the construction is repeated here stream for stream) as a family whose parameters are its
    canonical data (``q``, ``l``, ``u`` and the stored entries of ``P`` / ``A``).  Returns (family, cases): the original and the
    updated matrix entries and the expected x / objective of the four variants the reference test asserts
    (test_update_matrices.h: original, P updated, A updated, both; TESTS_TOL = 1e-4, duals zero)."""
    from numpy.random import Generator, PCG64
    from .ir import CanonFamily as _CF
    rg = Generator(PCG64(2))
    n, m, dens = 5, 8, 0.7
    A = sp.random(m, n, density=dens, format='csc', random_state=rg)
    P = sp.random(n, n, density=dens, random_state=rg)
    P = (P @ P.T).tocsc() + sp.eye(n, format='csc')
    Pu = sp.triu(P, format='csc'); Pu.sort_indices(); A.sort_indices()
    A_new = A.copy(); A_new.data = A_new.data + rg.standard_normal(A_new.nnz)
    Pu_new = Pu.copy(); Pu_new.data = Pu_new.data + 0.1 * rg.standard_normal(Pu_new.nnz)
    q = rg.standard_normal(n); l = -30 + rg.standard_normal(m); u = 30 + rg.standard_normal(m)
    fam = _CF.from_canonical_qp(name or 'osqp_update_matrices_5_8', Pu, q, A, l, u, n_eq=0, matrix_params=True)
    x_orig = np.array([-4.61725223e-01, 7.97298788e-01, 5.55470173e-04, 3.37603740e-01, -1.14060693e+00])
    x_pnew = np.array([-0.48845963, 0.70997599, -0.09017696, 0.33176037, -1.01867464])
    cases = dict(P=np.stack([Pu.data, Pu_new.data, Pu.data, Pu_new.data]), A=np.stack([A.data, A.data, A_new.data, A_new.data]),
                 x=np.stack([x_orig, x_pnew, x_orig, x_pnew]),
                 obj=np.array([-1.885431747787806, -1.7649689689774013, -1.8854317477878062, -1.764968968977401]))
    return fam, cases


def portfolio_qp(n=50, m=10, seed=0, name=None) -> CanonFamily:
    """The reference's portfolio test problem in its QP form (tests/test_E2E_QP.py:76-110, data :148-162; run there with OSQP):

        maximise a'w - |Sig_f_sqrt f|^2 - |d_sqrt . w|^2 - k_tc'|delta_w| + k_sh' min(0, w)
        s.t.     f = F'w,  1'w = 1,  |w|_1 <= L,  delta_w = w - w_prev

    canonical x = [w ; delta_w ; f ; t >= |delta_w| ; s <= min(0, w) ; v >= |w| ; g = Sig_f_sqrt f ; e = d_sqrt . w],
    P = 2 I on (g, e) -- the factors multiply variables inside sum_squares, so they enter the constraint matrix through
    auxiliaries like in cvxpy's DPP canonicalisation -- q = [-a ; 0 ; 0 ; k_tc ; -k_sh ; 0 ; 0 ; 0].
    rows: equalities  f - F'w = 0 (d0) | 1'w = 1 (d1) | delta_w - w = -w_prev (d3) | g - Sig f = 0 | e - d . w = 0
          inequalities  +-delta_w - t <= 0 | s <= 0 | s - w <= 0 | +-w - v <= 0 | 1'v <= L (d2)
    ``a``, ``w_prev``, ``k_tc``, ``k_sh``, ``L`` only touch q, l, u; ``F``, ``Sig_f_sqrt``, ``d_sqrt`` are matrix parameters."""
    rs = np.random.RandomState(seed)
    alpha = rs.randn(n)
    F = np.round(rs.randn(n, m))
    Sig = np.diag(rs.rand(m))
    dsq = rs.rand(n)
    params = _layout_params([('a', (n,), alpha), ('F', (n, m), F.flatten(order='F')), ('Sig_f_sqrt', (m, m), Sig.flatten(order='F')),
                             ('d_sqrt', (n,), dsq), ('k_tc', (n,), 0.01 * np.ones(n)), ('k_sh', (n,), 0.05 * np.ones(n)),
                             ('w_prev', (n,), np.zeros(n)), ('L', (), 1.6)])
    col = {p.name: p.col for p in params}
    n_theta = params[-1].col + params[-1].size + 1
    ow, od, of, ot, os_, ov, og, oe = 0, n, 2 * n, 2 * n + m, 3 * n + m, 4 * n + m, 5 * n + m, 5 * n + 2 * m
    nv = oe + n
    r_f, r_one, r_dw, r_g, r_e = 0, m, m + 1, m + 1 + n, 2 * m + 1 + n
    n_eq = r_e + n
    i_t1, i_t2, i_s1, i_s2, i_v1, i_v2, i_L = n_eq, n_eq + n, n_eq + 2 * n, n_eq + 3 * n, n_eq + 4 * n, n_eq + 5 * n, n_eq + 6 * n
    mt = i_L + 1
    n_ineq = mt - n_eq
    ent = []
    for j in range(m):
        ent.append((r_f + j, of + j, ('c', 1.0)))
        for i in range(n):
            ent.append((r_f + j, ow + i, ('p', col['F'] + i + n * j, -1.0)))          # -(F'w)_j = -sum_i F_ij w_i
        ent.append((r_g + j, og + j, ('c', 1.0)))
        for k in range(m):
            ent.append((r_g + j, of + k, ('p', col['Sig_f_sqrt'] + j + m * k, -1.0)))
    for i in range(n):
        ent += [(r_one, ow + i, ('c', 1.0)), (r_dw + i, od + i, ('c', 1.0)), (r_dw + i, ow + i, ('c', -1.0)),
                (r_e + i, oe + i, ('c', 1.0)), (r_e + i, ow + i, ('p', col['d_sqrt'] + i, -1.0)),
                (i_t1 + i, od + i, ('c', 1.0)), (i_t1 + i, ot + i, ('c', -1.0)),
                (i_t2 + i, od + i, ('c', -1.0)), (i_t2 + i, ot + i, ('c', -1.0)),
                (i_s1 + i, os_ + i, ('c', 1.0)), (i_s2 + i, os_ + i, ('c', 1.0)), (i_s2 + i, ow + i, ('c', -1.0)),
                (i_v1 + i, ow + i, ('c', 1.0)), (i_v1 + i, ov + i, ('c', -1.0)),
                (i_v2 + i, ow + i, ('c', -1.0)), (i_v2 + i, ov + i, ('c', -1.0)), (i_L, ov + i, ('c', 1.0))]
    ent.sort(key=lambda e: (e[1], e[0]))
    Ar = np.array([e[0] for e in ent]); Ac = np.array([e[1] for e in ent])
    indptr = np.zeros(nv + 1, dtype=np.int64)
    np.add.at(indptr, Ac + 1, 1); indptr = np.cumsum(indptr).astype(np.int32)
    mbA = _MapBuilder(len(ent), n_theta)
    for k, e in enumerate(ent):
        if e[2][0] == 'c':
            mbA.const(k, e[2][1])
        else:
            mbA.add(k, e[2][1], e[2][2])
    naux = m + n
    Pu = sp.csc_matrix((2.0 * np.ones(naux), (og + np.arange(naux), og + np.arange(naux))), shape=(nv, nv))
    mbP = _MapBuilder(naux, n_theta)
    for k in range(naux):
        mbP.const(k, 2.0)
    mq = _MapBuilder(nv, n_theta)
    for i in range(n):
        mq.add(ow + i, col['a'] + i, -1.0); mq.add(ot + i, col['k_tc'] + i, 1.0); mq.add(os_ + i, col['k_sh'] + i, -1.0)
    ml, mu = _MapBuilder(mt, n_theta), _MapBuilder(mt, n_theta)
    ml.const(r_one, 1.0); mu.const(r_one, 1.0)
    for i in range(n):
        ml.add(r_dw + i, col['w_prev'] + i, -1.0); mu.add(r_dw + i, col['w_prev'] + i, -1.0)
    for r in range(n_eq, mt):
        ml.const(r, -INF)
    mu.add(i_L, col['L'], 1.0)
    maps = {'A': mbA.csr(), 'P': mbP.csr(), 'q': mq.csr(), 'd': sp.csr_matrix((1, n_theta)), 'l': ml.csr(), 'u': mu.csr()}
    variables = [UserVar('w', (n,), ow + np.arange(n)), UserVar('delta_w', (n,), od + np.arange(n)), UserVar('f', (m,), of + np.arange(m))]
    duals = [UserDual('d0', 'y', (m,), r_f + np.arange(m)), UserDual('d1', 'y', (), np.array([r_one])),
             UserDual('d2', 'y', (), np.array([i_L])), UserDual('d3', 'y', (n,), r_dw + np.arange(n))]
    return CanonFamily(name or f'portfolio_qp_{n}_{m}', 'quadratic', nv, n_eq, n_ineq, params, maps,
                       {'P': _csc_pattern(Pu), 'A': (Ar.astype(np.int32), indptr, (mt, nv))}, variables, duals,
                       is_maximization=True)


def adp_socp(name=None) -> CanonFamily:
    """The reference's SOCP test problem (tests/test_E2E_SOCP.py:15-35, data :38-63; run there with ECOS among others):
    one step of approximate dynamic programming with two norm-bounded inputs,

        minimise |f + G u_0|^2 + |Rsqrt u_0|^2   s.t.  |u_i|_2 <= 0.1,  i = 0, 1        (u is 2 x 3, u_i its rows)

    ECOS form: x = [u(:) (Fortran order) ; t1 ; t2], minimise t1 + t2 with the two squared norms as rotated cones
    (1 + t, 1 - t, 2 v) in SOC and the two input bounds (0.1, u_i) in SOC(4); no LP cone, no equalities.
    ``f`` (what the current state enters through) is the per-instance vector; ``G`` and ``Rsqrt`` are matrix parameters."""
    n, m = 6, 3
    rs = np.random.RandomState(0)
    state = -2 * np.ones(6) + 4 * rs.rand(6)
    td = 0.1
    A = np.eye(6) + td * np.block([[np.zeros((3, 3)), np.eye(3)], [np.zeros((3, 3)), -np.diag(state[3:])]])
    Bm = td * np.vstack([np.zeros((3, 3)), np.diag(state[3:])])
    params = _layout_params([('Rsqrt', (m, m), np.sqrt(0.1) * np.ones(m)), ('f', (n,), A @ state), ('G', (n, m), Bm.flatten(order='F'))])
    col = {p.name: p.col for p in params}
    n_theta = params[-1].col + params[-1].size + 1
    nv = 2 * m + 2
    it1, it2 = 2 * m, 2 * m + 1
    u0 = lambda j: 2 * j                      # u[0, j] in Fortran order of the 2 x m variable
    u1 = lambda j: 2 * j + 1
    q = [n + 2, m + 2, m + 1, m + 1]
    rows = sum(q)
    ent, hmap = [], _MapBuilder(rows, n_theta)
    r = 0
    ent += [(r, it1, ('c', -1.0)), (r + 1, it1, ('c', 1.0))]; hmap.const(r, 1.0); hmap.const(r + 1, 1.0)
    for i in range(n):
        hmap.add(r + 2 + i, col['f'] + i, 2.0)
        for j in range(m):
            ent.append((r + 2 + i, u0(j), ('p', col['G'] + i + n * j, -2.0)))
    r += n + 2
    ent += [(r, it2, ('c', -1.0)), (r + 1, it2, ('c', 1.0))]; hmap.const(r, 1.0); hmap.const(r + 1, 1.0)
    for j in range(m):
        ent.append((r + 2 + j, u0(j), ('p', col['Rsqrt'] + j, -2.0)))
    r += m + 2
    for idx in (u0, u1):
        hmap.const(r, 0.1)
        for j in range(m):
            ent.append((r + 1 + j, idx(j), ('c', -1.0)))
        r += m + 1
    ent.sort(key=lambda e: (e[1], e[0]))
    Gr = np.array([e[0] for e in ent]); Gc = np.array([e[1] for e in ent])
    indptr = np.zeros(nv + 1, dtype=np.int64)
    np.add.at(indptr, Gc + 1, 1); indptr = np.cumsum(indptr).astype(np.int32)
    mg = _MapBuilder(len(ent), n_theta)
    for k, e in enumerate(ent):
        if e[2][0] == 'c':
            mg.const(k, e[2][1])
        else:
            mg.add(k, e[2][1], e[2][2])
    mc = _MapBuilder(nv, n_theta); mc.const(it1, 1.0); mc.const(it2, 1.0)
    maps = {'c': mc.csr(), 'd': sp.csr_matrix((1, n_theta)), 'A': sp.csr_matrix((0, n_theta)), 'b': sp.csr_matrix((0, n_theta)),
            'G': mg.csr(), 'h': hmap.csr()}
    variables = [UserVar('u', (2, m), np.arange(2 * m))]
    duals = [UserDual('d0', 'z', None, (n + 2) + (m + 2) + np.arange(2 * (m + 1)))]
    return CanonFamily(name or 'adp_socp_6_3', 'conic', nv, 0, rows, params, maps,
                       {'A': _csc_pattern(sp.csc_matrix((0, nv))), 'G': (Gr.astype(np.int32), indptr, (rows, nv))}, variables, duals,
                       is_maximization=False, cone_dims={'l': 0, 'q': q})


def network_lp(n=50, m=10, seed=0, name=None) -> CanonFamily:
    """The reference's network-flow LP (tests/test_E2E_LP.py:15-36, data :66-74; run there with solver='ECOS'):

        maximise w'f   s.t.  R f <= c,  f_min <= f <= f_max

    n flows over m links; ECOS form  min -w'f  s.t.  h - G f in R+^(m+2n),  G = [R ; -I ; I],  h = [c ; -f_min ; f_max],
    no equalities.  ``R`` (0/1 routing matrix, shared), ``c``, ``w``, ``f_min``, ``f_max`` are the user parameters; the
    vectors are the ones a batch varies.  Duals: d0 (link capacities), d1 (lower bounds), d2 (upper bounds)."""
    rs = np.random.RandomState(seed)
    R = np.round(rs.rand(m, n))
    cdef = n * (0.1 + 0.1 * rs.rand(m))
    wdef = rs.rand(n)
    Rs = sp.csc_matrix(R); Rs.eliminate_zeros()
    G = sp.vstack([Rs, -sp.identity(n), sp.identity(n)], format='csc'); G.sort_indices()
    A = sp.csc_matrix((0, n))
    params = _layout_params([('R', (m, n), R.flatten(order='F')), ('c', (m,), cdef), ('w', (n,), wdef),
                             ('f_min', (n,), np.zeros(n)), ('f_max', (n,), np.ones(n))])
    col = {p.name: p.col for p in params}
    n_theta = params[-1].col + params[-1].size + 1
    mt = m + 2 * n
    mc = _MapBuilder(n, n_theta)
    for i in range(n):
        mc.add(i, col['w'] + i, -1.0)
    mh = _MapBuilder(mt, n_theta)
    for i in range(m):
        mh.add(i, col['c'] + i, 1.0)
    for i in range(n):
        mh.add(m + i, col['f_min'] + i, -1.0)
        mh.add(m + n + i, col['f_max'] + i, 1.0)
    mg = _MapBuilder(G.nnz, n_theta)
    Gc = sp.coo_matrix(G)
    order = np.lexsort((Gc.row, Gc.col))
    for k, (r, c_, v) in enumerate(zip(Gc.row[order], Gc.col[order], Gc.data[order])):
        if r < m:
            mg.add(k, col['R'] + r + m * c_, 1.0)          # R is a dense parameter, Fortran order
        else:
            mg.const(k, v)
    maps = {'c': mc.csr(), 'd': sp.csr_matrix((1, n_theta)), 'A': sp.csr_matrix((0, n_theta)), 'b': sp.csr_matrix((0, n_theta)),
            'G': mg.csr(), 'h': mh.csr()}
    variables = [UserVar('f', (n,), np.arange(n))]
    duals = [UserDual('d0', 'z', (m,), np.arange(m)), UserDual('d1', 'z', (n,), m + np.arange(n)),
             UserDual('d2', 'z', (n,), m + n + np.arange(n))]
    return CanonFamily(name or f'network_lp_{n}_{m}', 'conic', n, 0, mt, params, maps,
                       {'A': _csc_pattern(A), 'G': _csc_pattern(G)}, variables, duals, is_maximization=True,
                       cone_dims={'l': mt, 'q': []})


def network_lp_batch(fam: CanonFamily, B: int, seed: int = 1):
    """Per-instance vectors drawn like the reference's assign_data (tests/test_E2E_LP.py:66-74)."""
    rng = np.random.default_rng(seed)
    n, m = fam.param('w').size, fam.param('c').size
    return {'c': n * (0.1 + 0.1 * rng.random((B, m))), 'w': rng.random((B, n)),
            'f_min': 0.05 * rng.random((B, n)), 'f_max': 1.0 + 0.1 * rng.random((B, n))}


def box_qp(n=6, m=8, seed=4, name=None) -> CanonFamily:
    """Small QP with two-sided constraints  l <= A x <= u  whose bounds are user parameters, built to exercise the
    per-instance corner cases of the path: a bound pair collapsing to an equality or opening to (-inf, inf) changes the
    constraint's TYPE (and hence rho_vec and the KKT factor: update_rho_vec, osqp_sources/src/auxil.c:100-142); two
    parallel rows with crossing bounds are primal infeasible; variable n-1 has no curvature and no constraint, so a
    non-zero cost on it is dual infeasible (unbounded).  User parameters: q (n), l (m), u (m)."""
    rs = np.random.RandomState(seed)
    Pd = np.concatenate([1.0 + rs.rand(n - 1), [0.0]])
    Pu = sp.csc_matrix(sp.diags(Pd))
    Ad = rs.randn(m, n) * (rs.rand(m, n) < 0.6)
    Ad[:, n - 1] = 0.0
    Ad[1] = Ad[0]                                   # rows 0 and 1 are parallel
    for i in range(m):
        if not Ad[i].any():
            Ad[i, rs.randint(n - 1)] = 1.0
    A = sp.csc_matrix(Ad); A.sort_indices()
    x0 = rs.randn(n)
    q0 = np.concatenate([rs.randn(n - 1), [0.0]])
    l0 = Ad @ x0 - 0.5 - rs.rand(m); u0 = Ad @ x0 + 0.5 + rs.rand(m)
    params = _layout_params([('q', (n,), q0), ('l', (m,), l0), ('u', (m,), u0)])
    n_theta = n + 2 * m + 1
    maps = {}
    mb = _MapBuilder(Pu.nnz, n_theta)
    Pu.sort_indices()
    for k, v in enumerate(Pu.data):
        mb.const(k, v)
    maps['P'] = mb.csr()
    mb = _MapBuilder(A.nnz, n_theta)
    for k, v in enumerate(A.data):
        mb.const(k, v)
    maps['A'] = mb.csr()
    mq, ml, mu = _MapBuilder(n, n_theta), _MapBuilder(m, n_theta), _MapBuilder(m, n_theta)
    for i in range(n):
        mq.add(i, params[0].col + i, 1.0)
    for j in range(m):
        ml.add(j, params[1].col + j, 1.0); mu.add(j, params[2].col + j, 1.0)
    maps['q'], maps['l'], maps['u'] = mq.csr(), ml.csr(), mu.csr()
    maps['d'] = sp.csr_matrix((1, n_theta))
    return CanonFamily(name or f'box_qp_{n}_{m}', 'quadratic', n, 0, m, params, maps,
                       {'P': _csc_pattern(Pu), 'A': _csc_pattern(A)}, [UserVar('x', (n,), np.arange(n))],
                       [UserDual('d0', 'y', (m,), np.arange(m))])
