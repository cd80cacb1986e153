"""Modified Ruiz equilibration of the QP data, as OSQP applies it at setup.

Reference behaviour being matched (a10): osqp_sources/src/scaling.c:44-156
(`scale_data`), with `limit_scaling` :7-14 and the KKT column norms of :28-42.
The result (D, E, c and the scaled P, A) must equal the reference's bit for bit
where possible, because the ADMM trajectory -- and therefore the iterate at which
the eps 1e-3 stopping test fires -- depends on it (last paragraph).
Sparse formulation: only stored entries are touched.
"""
import numpy as np
import scipy.sparse as sp

MIN_SCALING = 1e-4
MAX_SCALING = 1e4


def _clamp_norms(v):
    v = np.array(v, dtype=float, copy=True)
    v[v < MIN_SCALING] = 1.0
    v[v > MAX_SCALING] = MAX_SCALING
    return v


def _sequential_sum(v):
    acc = 0.0
    for t in v.tolist():
        acc += t
    return acc


def _col_absmax(M: sp.csc_matrix, n_cols):
    out = np.zeros(n_cols)
    if M.nnz:
        np.maximum.at(out, np.repeat(np.arange(n_cols), np.diff(M.indptr)), np.abs(M.data))
    return out


def _row_absmax(M: sp.csc_matrix, n_rows):
    out = np.zeros(n_rows)
    if M.nnz:
        np.maximum.at(out, M.indices, np.abs(M.data))
    return out


def ruiz_equilibrate(P_upper: sp.csc_matrix, A: sp.csc_matrix, q: np.ndarray, n_iter: int = 10):
    """Returns dict(P=scaled upper-tri CSC, A=scaled CSC, q=scaled q, D, E, c).

    P_upper holds the upper triangle only (OSQP convention); symmetric column norms
    are therefore max(column norm, row norm) of the stored triangle."""
    P = sp.csc_matrix(P_upper, dtype=float, copy=True)
    A = sp.csc_matrix(A, dtype=float, copy=True)
    n, m = P.shape[0], A.shape[0]
    q = np.array(q, dtype=float, copy=True)
    D, E, c = np.ones(n), np.ones(m), 1.0
    p_cols = np.repeat(np.arange(n), np.diff(P.indptr))
    a_cols = np.repeat(np.arange(n), np.diff(A.indptr))
    for _ in range(int(n_iter)):
        norm_P = np.maximum(_col_absmax(P, n), _row_absmax(P, n))
        d = 1.0 / np.sqrt(_clamp_norms(np.maximum(norm_P, _col_absmax(A, n))))
        e = 1.0 / np.sqrt(_clamp_norms(_row_absmax(A, m)))
        P.data = (d[P.indices] * P.data) * d[p_cols]          # P <- D P D  (pre-mult then post-mult)
        A.data = (e[A.indices] * A.data) * d[a_cols]          # A <- E A D
        q = d * q
        D = D * d
        E = E * e
        # cost normalisation
        norm_P = np.maximum(_col_absmax(P, n), _row_absmax(P, n))
        mean_P = _sequential_sum(norm_P) / n if n else 0.0   # left-to-right like vec_mean (lin_alg.c)
        norm_q = float(_clamp_norms([np.abs(q).max() if n else 0.0])[0])
        gamma = 1.0 / float(_clamp_norms([max(mean_P, norm_q)])[0])
        P.data = P.data * gamma
        q = q * gamma
        c = c * gamma
    return dict(P=P, A=A, q=q, D=D, E=E, c=c)
