"""KKT assembly, fill-reducing ordering, and LDL' factorisation for the ADMM linear system.

What the reference does at setup (a10): `set_rho_vec` (osqp_sources/src/auxil.c:76-98),
`form_KKT` (src/kkt.c:6-177), AMD ordering + `QDLDL_etree`/`QDLDL_factor`
(lin_sys/direct/qdldl/qdldl_interface.c:53-173, qdldl_sources/src/qdldl.c:11-233).

What is different here, on purpose: the factor is not going to be used by a scalar
column-by-column triangular solve but by 32 lanes of a warp, so the ordering is chosen
for *parallel depth* as well as fill -- a minimum-degree ordering whose elimination tree
is then re-sequenced level by level (leaves first).  Any topological re-sequencing of
an elimination tree leaves the fill unchanged, so nnz(L) is that of the minimum-degree
ordering while all columns of one level are mutually independent.  The KKT solve is
exact whatever the ordering, so the ADMM iterates are unchanged up to rounding.
"""
from dataclasses import dataclass
from typing import List

import numpy as np
import scipy.sparse as sp

OSQP_INFTY = 1e30
MIN_SCALING = 1e-4
RHO_MIN, RHO_MAX = 1e-6, 1e6
RHO_TOL = 1e-4
RHO_EQ_OVER_RHO_INEQ = 1e3

CONSTR_LOOSE, CONSTR_INEQ, CONSTR_EQ = -1, 0, 1


def constraint_types(l_scaled: np.ndarray, u_scaled: np.ndarray) -> np.ndarray:
    """-1 loose / 0 inequality / +1 equality, decided on the SCALED bounds like the reference."""
    loose = (l_scaled < -OSQP_INFTY * MIN_SCALING) & (u_scaled > OSQP_INFTY * MIN_SCALING)
    eq = ~loose & (u_scaled - l_scaled < RHO_TOL)
    return np.where(loose, CONSTR_LOOSE, np.where(eq, CONSTR_EQ, CONSTR_INEQ)).astype(np.int32)


def rho_vector(ctype: np.ndarray, rho: float) -> np.ndarray:
    rho = min(max(rho, RHO_MIN), RHO_MAX)
    return np.where(ctype == CONSTR_LOOSE, RHO_MIN,
                    np.where(ctype == CONSTR_EQ, RHO_EQ_OVER_RHO_INEQ * rho, rho)).astype(float)


def assemble_kkt(P_upper: sp.csc_matrix, A: sp.csc_matrix, sigma: float, rho_vec: np.ndarray) -> sp.csc_matrix:
    """Full symmetric K = [[P + sigma I, A'], [A, -diag(1/rho_vec)]] as CSC."""
    n, m = P_upper.shape[0], A.shape[0]
    Pfull = P_upper + sp.triu(P_upper, 1).T + sigma * sp.eye(n)
    K = sp.bmat([[Pfull, A.T], [A, -sp.diags(1.0 / rho_vec)]], format='csc')
    K.sort_indices()
    return K


def minimum_degree_order(pattern: sp.spmatrix) -> np.ndarray:
    """Exact minimum-degree elimination order on the graph of a symmetric pattern.
    Ties broken by smallest index.  Sizes here are a few hundred to a few thousand
    nodes, so the plain set-based algorithm is adequate (it runs once per family)."""
    n = pattern.shape[0]
    S = sp.csr_matrix(pattern)
    adj = [set(S.indices[S.indptr[i]:S.indptr[i + 1]].tolist()) - {i} for i in range(n)]
    alive = np.ones(n, dtype=bool)
    deg = np.array([len(a) for a in adj])
    order = []
    big = n + 1
    for _ in range(n):
        d = np.where(alive, deg, big)
        v = int(np.argmin(d))
        order.append(v)
        alive[v] = False
        nb = adj[v]
        for a in nb:
            adj[a].discard(v)
        nbl = list(nb)
        for a in nbl:
            adj[a] |= nb
            adj[a].discard(a)
            deg[a] = len(adj[a])
        adj[v] = set()
    return np.array(order, dtype=np.int64)


def _symbolic(pattern: sp.spmatrix, order: np.ndarray):
    """Column structures of L (strictly lower, in pivot positions) and the etree parent."""
    n = pattern.shape[0]
    inv = np.empty(n, dtype=np.int64); inv[order] = np.arange(n)
    S = sp.csr_matrix(pattern)
    # adjacency in pivot positions, only towards higher positions
    higher = [set() for _ in range(n)]
    for i in range(n):
        pi = inv[i]
        for j in S.indices[S.indptr[i]:S.indptr[i + 1]]:
            pj = inv[j]
            if pj > pi:
                higher[pi].add(int(pj))
    parent = -np.ones(n, dtype=np.int64)
    struct: List[np.ndarray] = [None] * n
    for k in range(n):
        s = higher[k]
        if s:
            p = min(s)
            parent[k] = p
            higher[p] |= (s - {p})
        struct[k] = np.array(sorted(s), dtype=np.int64)
    return struct, parent


def level_resequence(struct, parent) -> np.ndarray:
    """Positions -> new positions such that etree levels (leaves = 0) are contiguous."""
    n = len(struct)
    level = np.zeros(n, dtype=np.int64)
    for k in range(n):                       # children have smaller positions than parents
        p = parent[k]
        if p >= 0:
            level[p] = max(level[p], level[k] + 1)
    new_order = np.lexsort((np.arange(n), level))   # stable by level then old position
    return new_order, level


@dataclass
class LDLFactor:
    perm: np.ndarray        # perm[k] = original KKT index eliminated k-th
    level: np.ndarray       # etree level of position k (non-decreasing in k)
    L: np.ndarray           # dense unit-lower factor in pivot positions (strictly lower part stored, unit diag implied)
    D: np.ndarray           # diagonal
    Lpattern: np.ndarray    # bool, symbolic strictly-lower pattern (superset of numerically nonzero entries)
    n_pos: int              # number of positive pivots (must equal n_var for a convex QP)

    @property
    def nnz(self):
        return int(self.Lpattern.sum())


def dense_ldl(Kp: np.ndarray, pattern: np.ndarray):
    """Right-looking LDL' without pivoting on the permuted dense matrix (quasi-definite => stable
    for any order).  Only entries inside the symbolic pattern are ever non-zero."""
    n = Kp.shape[0]
    S = np.array(Kp, dtype=float, copy=True)
    L = np.zeros((n, n))
    D = np.zeros(n)
    for k in range(n):
        D[k] = S[k, k]
        if D[k] == 0.0:
            raise ZeroDivisionError('zero pivot in LDL factorisation')
        rows = np.nonzero(pattern[:, k])[0]
        if rows.size:
            col = S[rows, k]
            lk = col / D[k]
            L[rows, k] = lk
            S[np.ix_(rows, rows)] -= np.outer(lk, col)
    return L, D


def factorize(K: sp.csc_matrix, n_var: int, pattern: sp.spmatrix = None) -> LDLFactor:
    """pattern: optional structural pattern (superset of K's nonzeros) to order and analyse instead of K's own."""
    n = K.shape[0]
    Ks = K if pattern is None else sp.csc_matrix(pattern)
    patt = sp.csr_matrix((np.ones(Ks.nnz), Ks.indices, Ks.indptr), shape=K.shape)
    md = minimum_degree_order(patt)
    struct, parent = _symbolic(patt, md)
    reseq, _ = level_resequence(struct, parent)
    perm = md[reseq]
    struct, parent = _symbolic(patt, perm)
    _, level = level_resequence(struct, parent)
    assert np.all(np.diff(level) >= 0)
    pattern = np.zeros((n, n), dtype=bool)
    for k in range(n):
        pattern[struct[k], k] = True
    Kp = K.toarray()[np.ix_(perm, perm)]
    L, D = dense_ldl(Kp, pattern)
    n_pos = int((D > 0).sum())
    if n_pos != n_var:
        raise ValueError('KKT matrix has the wrong inertia: the problem seems to be non-convex '
                         '(reference: qdldl_interface.c:93-99)')
    return LDLFactor(perm=perm, level=level, L=L, D=D, Lpattern=pattern, n_pos=n_pos)
