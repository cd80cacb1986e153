"""oracle/grad_numpy.py -- TEST INFRASTRUCTURE (oracle), not product code.

numpy restatement of the reference's QP backward pass (`gradient=True`):
  cpg_osqp_gradient()      cvxpygen/templates/cpg_osqp_grad_compute.c.jinja2:432-531
  K / K_true definition    cvxpygen/writer.py:354-371   (P + 1e-6 I, -1e-6 I regularisation; K_true exact)
  un-canonicalisation      cvxpygen/writer.py:268-303   (dp = sum_id map_id' d(id))

Given a canonical solution (x, y) of  min 1/2 x'Px + q'x  s.t. l <= Ax <= u  and an upstream gradient dx on x:
  active set   a_i = -1 if y_i < -1e-12, +1 if y_i > 1e-12, else 0                       (:437-454)
  solve        K r = [dx; 0],   K = [[P + 1e-6 I, A_act'], [A_act, -1e-6 I]]  (inactive rows: pivot -1, r_i = 0)
  refine x3    r += K^{-1} ([dx;0] - K_true r),  K_true = [[P, A_act'],[A_act, 0]]      (:456-490)
  dq = -r_x ;  dl_i = r_{n+i} (a_i = -1) ;  du_i = r_{n+i} (a_i = +1)                    (:492-511)
  dP_ij = -1/2 (r_i x_j + x_i r_j) ;  dA_ij = -(r_{n+i} x_j + y_i r_j) on active rows    (:513-529)
The reference keeps one LDL' factor alive and up/down-dates it between calls; mathematically each call solves the
system above, which is what is restated here with a dense factorisation per instance.

Pinned by tests/test_grad_oracle.py against the reference's own generated C (oracle/_ref/libgrad_ref_*.so, built by
oracle/build_grad_ref.py from the reference's templates) and its golden vectors, and against finite differences.
"""
import numpy as np
import scipy.sparse as sp

ACTIVE_TOL = 1e-12
REG = 1e-6


def qp_backward(P, A, x, y, dx, n_refine=3):
    """P (n,n) full symmetric or upper, A (m,n); x,y,dx batches (B,n),(B,m),(B,n). Returns dq, dl, du, r."""
    P = sp.csc_matrix(P)
    Pf = (sp.triu(P) + sp.triu(P, 1).T).toarray()
    Ad = sp.csc_matrix(A).toarray()
    n, m = Pf.shape[0], Ad.shape[0]
    x = np.atleast_2d(x); y = np.atleast_2d(y); dx = np.atleast_2d(dx)
    B = x.shape[0]
    dq = np.zeros((B, n)); dl = np.zeros((B, m)); du = np.zeros((B, m)); R = np.zeros((B, n + m))
    for b in range(B):
        a = np.where(y[b] < -ACTIVE_TOL, -1, np.where(y[b] > ACTIVE_TOL, 1, 0))
        act = a != 0
        Aa = Ad * act[:, None]
        K = np.zeros((n + m, n + m))
        K[:n, :n] = Pf + REG * np.eye(n)
        K[:n, n:] = Aa.T
        K[n:, :n] = Aa
        K[n:, n:] = np.diag(np.where(act, -REG, -1.0))
        Kt = np.zeros_like(K)
        Kt[:n, :n] = Pf; Kt[:n, n:] = Aa.T; Kt[n:, :n] = Aa
        rhs = np.concatenate([dx[b], np.zeros(m)])
        r = np.linalg.solve(K, rhs)
        for _ in range(n_refine):
            delta = rhs - Kt @ r
            delta[n:][~act] = 0.0
            r = r + np.linalg.solve(K, delta)
        R[b] = r
        dq[b] = -r[:n]
        dl[b] = np.where(a == -1, r[n:], 0.0)
        du[b] = np.where(a == 1, r[n:], 0.0)
    return dq, dl, du, R


def param_gradient(fam, dq, dl, du, names=None):
    """dtheta for vector-valued canonical ids: sum of map' d(id) restricted to the given user parameters."""
    cols = fam.param_columns(names)
    out = np.zeros((dq.shape[0], len(cols)))
    for pid, d in (('q', dq), ('l', dl), ('u', du)):
        M = fam.maps.get(pid)
        if M is not None and M.nnz:
            out += d @ M[:, cols].toarray()
    return out


def qp_backward_mat(P_pattern, A_pattern, Px, Ax, x, y, dx, n_refine=3):
    """Per-instance matrices (SURVEY rows a16 + f2).  P_pattern / A_pattern: (indices, indptr, shape) CSC (P upper triangle);
    Px (B, nnzP), Ax (B, nnzA) the instances' entries.  Returns dq, dl, du, dP (B, nnzP), dA (B, nnzA):
      dP_k = -1/2 (r_i x_j + x_i r_j),  dA_k = -(r_{n+i} x_j + y_i r_j) on active rows, 0 otherwise
    for the stored entry k = (i, j)   (cpg_osqp_grad_compute.c.jinja2:513-529)."""
    Pi, Pp, Ps = P_pattern; Ai, Ap, As = A_pattern
    n, m = Ps[0], As[0]
    Prow = np.asarray(Pi); Pcol = np.repeat(np.arange(n), np.diff(Pp))
    Arow = np.asarray(Ai); Acol = np.repeat(np.arange(n), np.diff(Ap))
    x = np.atleast_2d(x); y = np.atleast_2d(y); dx = np.atleast_2d(dx)
    B = x.shape[0]
    dq = np.zeros((B, n)); dl = np.zeros((B, m)); du = np.zeros((B, m))
    dP = np.zeros((B, len(Prow))); dA = np.zeros((B, len(Arow)))
    for b in range(B):
        P = sp.csc_matrix((Px[b], Pi, Pp), shape=Ps); A = sp.csc_matrix((Ax[b], Ai, Ap), shape=As)
        q_, l_, u_, R = qp_backward(P, A, x[b], y[b], dx[b], n_refine)
        r = R[0]
        dq[b], dl[b], du[b] = q_[0], l_[0], u_[0]
        dP[b] = -0.5 * (r[Prow] * x[b, Pcol] + x[b, Prow] * r[Pcol])
        act = np.abs(y[b]) > ACTIVE_TOL
        dA[b] = np.where(act[Arow], -(r[n + Arow] * x[b, Acol] + y[b, Arow] * r[Acol]), 0.0)
    return dq, dl, du, dP, dA


def param_gradient_mat(fam, dq, dl, du, dP, dA, names=None):
    """dtheta including the matrix ids: sum over id in {q, l, u, P, A} of map_id' d(id)   (cvxpygen/writer.py:268-303)."""
    cols = fam.param_columns(names)
    out = param_gradient(fam, dq, dl, du, names)
    for pid, d in (('P', dP), ('A', dA)):
        M = fam.maps.get(pid)
        if M is not None and M.nnz:
            out += d @ M[:, cols].toarray()
    return out
