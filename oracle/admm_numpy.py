"""oracle/admm_numpy.py -- TEST INFRASTRUCTURE (oracle), not product code.

numpy restatement of the reference's QP hot path (OSQP 0.6.2 ADMM as driven by
cvxpygen's generated cpg_solve), vectorised over a batch of instances that share
P and A and differ in q, l, u.  Each step cites the reference lines it restates;
paths are relative to cvxpygen/solvers/osqp-python/osqp_sources/.

Parity pinning: tests/test_oracle.py checks this file against
  * OSQP's own known-answer tests (tests/basic_qp, basic_qp2, primal_infeasibility,
    primal_dual_infeasibility, unconstrained generate_problem.py values), and
  * the unmodified reference compiled into oracle/_ref/libosqp_ref.so, and
  * the golden vectors in tests/golden/ that were generated with oracle/_ref.

The linear algebra differs from the reference on purpose (dense inverse of the KKT
matrix instead of AMD + QDLDL): the KKT solve is exact in both, so the iterates
agree to rounding (~1e-12), which is what the tests assert.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.
"""
import numpy as np
import scipy.sparse as sp

OSQP_INFTY = 1e30          # include/constants.h:100
MIN_SCALING, MAX_SCALING = 1e-4, 1e4   # constants.h:87-88
RHO_MIN, RHO_MAX = 1e-6, 1e6           # constants.h:69-70
RHO_TOL = 1e-4                          # constants.h:72
RHO_EQ_OVER_RHO_INEQ = 1e3              # constants.h:71
DIVISION_TOL = 1.0 / OSQP_INFTY         # constants.h:104

SOLVED, SOLVED_INACCURATE = 1, 2
PRIMAL_INFEASIBLE, PRIMAL_INFEASIBLE_INACCURATE = -3, 3
DUAL_INFEASIBLE, DUAL_INFEASIBLE_INACCURATE = -4, 4
MAX_ITER_REACHED, NON_CVX, UNSOLVED = -2, -7, -10   # constants.h:18-30

DEFAULTS = dict(rho=0.1, sigma=1e-6, alpha=1.6, scaling=10, adaptive_rho=1, adaptive_rho_interval=0,
                adaptive_rho_tolerance=5.0, max_iter=4000, eps_abs=1e-3, eps_rel=1e-3,
                eps_prim_inf=1e-4, eps_dual_inf=1e-4, scaled_termination=0, check_termination=25)
# cvxpygen's OSQP settings table: cvxpygen/solvers/osqp.py:102-115 (on OSQP defaults constants.h:59-85)


def _limit_scaling(v):
    """scaling.c:7-14"""
    v = np.where(v < MIN_SCALING, 1.0, v)
    return np.where(v > MAX_SCALING, MAX_SCALING, v)


def ruiz_scale(P_full, A, q, n_iter):
    """scale_data, scaling.c:44-156.  P_full symmetric dense, A dense.  Returns scaled P,A,q and D,E,c."""
    n, m = P_full.shape[0], A.shape[0]
    P, A, q = P_full.copy(), A.copy(), q.copy()
    D, E, c = np.ones(n), np.ones(m), 1.0
    for _ in range(n_iter):
        # compute_inf_norm_cols_KKT, scaling.c:28-42
        Dt = np.maximum(np.abs(P).max(axis=0) if n else np.zeros(0),
                        np.abs(A).max(axis=0) if m else np.zeros(n))
        Et = np.abs(A).max(axis=1) if m else np.zeros(0)
        Dt = 1.0 / np.sqrt(_limit_scaling(Dt))
        Et = 1.0 / np.sqrt(_limit_scaling(Et))
        P = Dt[:, None] * P * Dt[None, :]
        A = Et[:, None] * A * Dt[None, :]
        q = Dt * q
        D, E = D * Dt, E * Et
        # cost normalisation, scaling.c:112-142
        c_temp = np.abs(P).max(axis=0).mean()
        inf_norm_q = float(_limit_scaling(np.array([np.abs(q).max() if n else 0.0]))[0])
        c_temp = max(c_temp, inf_norm_q)
        c_temp = 1.0 / float(_limit_scaling(np.array([c_temp]))[0])
        P, q, c = P * c_temp, q * c_temp, c * c_temp
    return P, A, q, D, E, c


class AdmmOracle:
    """setup once per family (osqp_setup, src/osqp.c:76-283), then solve batches."""

    def __init__(self, P, q, A, l, u, **settings):
        s = dict(DEFAULTS); s.update(settings)
        self.s = s
        P = sp.csc_matrix(P)
        Pu = sp.triu(P)
        self.P0 = (Pu + sp.triu(Pu, 1).T).toarray()       # full symmetric from the upper triangle
        self.A0 = sp.csc_matrix(A).toarray()
        self.n, self.m = self.P0.shape[0], self.A0.shape[0]
        self.q0 = np.asarray(q, float).copy()
        self.l0 = np.clip(np.asarray(l, float), -OSQP_INFTY, OSQP_INFTY)
        self.u0 = np.clip(np.asarray(u, float), -OSQP_INFTY, OSQP_INFTY)
        if s['scaling']:
            self.P, self.A, _, self.D, self.E, self.c = ruiz_scale(self.P0, self.A0, self.q0, int(s['scaling']))
        else:
            self.P, self.A = self.P0.copy(), self.A0.copy()
            self.D, self.E, self.c = np.ones(self.n), np.ones(self.m), 1.0
        self.Dinv, self.Einv, self.cinv = 1.0 / self.D, 1.0 / self.E, 1.0 / self.c
        # osqp.c:267-279: no PROFILING timer => interval = 4 * check_termination (or 100)
        if s['adaptive_rho'] and not s['adaptive_rho_interval']:
            self.s['adaptive_rho_interval'] = 4 * s['check_termination'] if s['check_termination'] else 100
        self._kinv_cache = {}

    # -- rho_vec / constraint types: set_rho_vec + update_rho_vec, auxil.c:76-142
    def _rho_vec(self, l, u, rho):
        loose = (l < -OSQP_INFTY * MIN_SCALING) & (u > OSQP_INFTY * MIN_SCALING)
        eq = ~loose & (u - l < RHO_TOL)
        rv = np.where(loose, RHO_MIN, np.where(eq, RHO_EQ_OVER_RHO_INEQ * rho[:, None], rho[:, None]))
        return rv

    def _kkt_inverse(self, rho_vec):
        """K = [[P + sigma I, A'],[A, -diag(1/rho_vec)]] (form_KKT, src/kkt.c:6-177); dense inverse
        stands in for permute_KKT + QDLDL_factor (qdldl_interface.c:53-173) + QDLDL_solve (qdldl.c:269)."""
        key = rho_vec.tobytes()
        Ki = self._kinv_cache.get(key)
        if Ki is None:
            n, m = self.n, self.m
            K = np.zeros((n + m, n + m))
            K[:n, :n] = self.P + self.s['sigma'] * np.eye(n)
            K[:n, n:] = self.A.T
            K[n:, :n] = self.A
            K[n:, n:] = -np.diag(1.0 / rho_vec)
            Ki = np.linalg.inv(K)
            if len(self._kinv_cache) > 64:
                self._kinv_cache.clear()
            self._kinv_cache[key] = Ki
        return Ki

    def solve_batch(self, q=None, l=None, u=None, B=None, x0=None, y0=None):
        s, n, m = self.s, self.n, self.m
        for a in (q, l, u):
            if a is not None:
                B = np.asarray(a).shape[0]
        B = 1 if B is None else B
        bc = lambda a, d: np.tile(d, (B, 1)) if a is None else np.array(a, dtype=float, copy=True).reshape(B, -1)
        q = bc(q, self.q0)
        l = np.clip(bc(l, self.l0), -OSQP_INFTY, OSQP_INFTY)
        u = np.clip(bc(u, self.u0), -OSQP_INFTY, OSQP_INFTY)
        # osqp_update_lin_cost osqp.c:752-781 / osqp_update_bounds :784-827 (and scale_data :151-153)
        q = self.c * self.D * q
        l, u = self.E * l, self.E * u
        rho = np.full(B, min(max(s['rho'], RHO_MIN), RHO_MAX))
        rho_vec = self._rho_vec(l, u, rho)
        P, A, D, E, Dinv, Einv, c, cinv = self.P, self.A, self.D, self.E, self.Dinv, self.Einv, self.c, self.cinv
        sigma, alpha = s['sigma'], s['alpha']
        # cold_start auxil.c:155-159 / osqp_warm_start osqp.c:929-958 (x<-Dinv x, y<-c Einv y, z<-Ax)
        if x0 is not None and y0 is not None:
            x = np.asarray(x0, float).reshape(B, n) * Dinv
            y = np.asarray(y0, float).reshape(B, m) * Einv * c
            z = x @ A.T
        else:
            x, z, y = np.zeros((B, n)), np.zeros((B, m)), np.zeros((B, m))
        status = np.full(B, UNSOLVED)
        it_out = np.zeros(B, dtype=np.int32)
        pri_res, dua_res = np.zeros(B), np.zeros(B)
        obj = np.zeros(B)
        rho_updates = np.zeros(B, dtype=np.int32)
        active = np.arange(B)
        dx = np.zeros((B, n)); dy = np.zeros((B, m))
        ct, ari = s['check_termination'], s['adaptive_rho_interval']

        def info(ix):
            """update_info auxil.c:564-629 -> compute_pri_res :240-254, compute_dua_res :287-318"""
            Ax = x[ix] @ A.T
            rp = Ax - z[ix]
            Px = x[ix] @ P
            Aty = y[ix] @ A
            rd = q[ix] + Px + Aty
            if s['scaling'] and not s['scaled_termination']:
                pr = np.abs(Einv * rp).max(axis=1) if m else np.zeros(len(ix))
                dr = cinv * np.abs(Dinv * rd).max(axis=1)
            else:
                pr = np.abs(rp).max(axis=1) if m else np.zeros(len(ix))
                dr = np.abs(rd).max(axis=1)
            return Ax, Px, Aty, rp, rd, pr, dr

        def check(ix, Ax, Px, Aty, pr, dr, approximate):
            """check_termination auxil.c:681-786; returns status codes (UNSOLVED = keep going)"""
            k = len(ix)
            out = np.full(k, UNSOLVED)
            ea, er, epi, edi = s['eps_abs'], s['eps_rel'], s['eps_prim_inf'], s['eps_dual_inf']
            if approximate:
                ea, er, epi, edi = 10 * ea, 10 * er, 10 * epi, 10 * edi
            noncvx = (pr > OSQP_INFTY) | (dr > OSQP_INFTY)
            unscale = s['scaling'] and not s['scaled_termination']
            # compute_pri_tol :256-285 / compute_dua_tol :320-359
            if m:
                if unscale:
                    mp = np.maximum(np.abs(Einv * z[ix]).max(axis=1), np.abs(Einv * Ax).max(axis=1))
                else:
                    mp = np.maximum(np.abs(z[ix]).max(axis=1), np.abs(Ax).max(axis=1))
                eps_prim = ea + er * mp
                prim_ok = pr < eps_prim
            else:
                prim_ok = np.ones(k, bool)
            if unscale:
                md = cinv * np.maximum(np.maximum(np.abs(Dinv * q[ix]).max(axis=1), np.abs(Dinv * Aty).max(axis=1)),
                                       np.abs(Dinv * Px).max(axis=1))
            else:
                md = np.maximum(np.maximum(np.abs(q[ix]).max(axis=1), np.abs(Aty).max(axis=1)), np.abs(Px).max(axis=1))
            eps_dual = ea + er * md
            dual_ok = dr < eps_dual
            prim_inf = np.zeros(k, bool); dual_inf = np.zeros(k, bool)
            # is_primal_infeasible :361-424
            for j in np.nonzero(~prim_ok)[0]:
                i = ix[j]
                d = dy[i].copy()
                up_inf = u[i] > OSQP_INFTY * MIN_SCALING
                lo_inf = l[i] < -OSQP_INFTY * MIN_SCALING
                d = np.where(up_inf & lo_inf, 0.0, np.where(up_inf, np.minimum(d, 0.0), np.where(lo_inf, np.maximum(d, 0.0), d)))
                nd = np.abs(E * d).max() if unscale else np.abs(d).max()
                if nd > DIVISION_TOL:
                    lhs = (u[i] * np.maximum(d, 0) + l[i] * np.minimum(d, 0)).sum()
                    if lhs < epi * nd:
                        Atd = d @ A
                        if unscale:
                            Atd = Dinv * Atd
                        prim_inf[j] = np.abs(Atd).max() < epi * nd
            # is_dual_infeasible :426-512
            for j in np.nonzero(~dual_ok)[0]:
                i = ix[j]
                d = dx[i]
                if unscale:
                    nd, cs = np.abs(D * d).max(), c
                else:
                    nd, cs = np.abs(d).max(), 1.0
                if nd > DIVISION_TOL and (q[i] @ d) < cs * edi * nd:
                    Pd = P @ d
                    if unscale:
                        Pd = Dinv * Pd
                    if np.abs(Pd).max() < cs * edi * nd:
                        Ad = A @ d
                        if unscale:
                            Ad = Einv * Ad
                        bad = ((u[i] < OSQP_INFTY * MIN_SCALING) & (Ad > edi * nd)) | \
                              ((l[i] > -OSQP_INFTY * MIN_SCALING) & (Ad < -edi * nd))
                        dual_inf[j] = not bad.any()
            solved = prim_ok & dual_ok
            out[dual_inf] = DUAL_INFEASIBLE_INACCURATE if approximate else DUAL_INFEASIBLE
            out[prim_inf] = PRIMAL_INFEASIBLE_INACCURATE if approximate else PRIMAL_INFEASIBLE
            out[solved] = SOLVED_INACCURATE if approximate else SOLVED
            out[noncvx] = NON_CVX
            return out

        last_checked = np.zeros(B, bool)
        it = 0
        for it in range(1, s['max_iter'] + 1):        # osqp.c:354
            if active.size == 0:
                it -= 1
                break
            ix = active
            x_prev, z_prev = x[ix].copy(), z[ix].copy()
            riv = 1.0 / rho_vec[ix]
            # compute_rhs auxil.c:161-175
            rhs = np.concatenate([sigma * x_prev - q[ix], z_prev - riv * y[ix]], axis=1)
            # solve_linsys_qdldl qdldl_interface.c:350-376
            sol = np.empty_like(rhs)
            keys = {}
            for j, i in enumerate(ix):
                keys.setdefault(rho_vec[i].tobytes(), []).append(j)
            for key, js in keys.items():
                Ki = self._kkt_inverse(rho_vec[ix[js[0]]])
                sol[js] = rhs[js] @ Ki.T
            xt = sol[:, :n]
            zt = rhs[:, n:] + riv * sol[:, n:]
            # update_x :185-198, update_z :200-212 (+ project, proj.c:4-14), update_y :214-225
            xn = alpha * xt + (1.0 - alpha) * x_prev
            dx[ix] = xn - x_prev
            zn = alpha * zt + (1.0 - alpha) * z_prev + riv * y[ix]
            zn = np.minimum(np.maximum(zn, l[ix]), u[ix])
            dyn = rho_vec[ix] * (alpha * zt + (1.0 - alpha) * z_prev - zn)
            dy[ix] = dyn
            x[ix], z[ix], y[ix] = xn, zn, y[ix] + dyn
            can_check = bool(ct) and it % ct == 0
            last_checked[:] = False
            have_info = False
            if can_check:                               # osqp.c:411-444
                Ax, Px, Aty, rp, rd, pr, dr = info(ix)
                have_info = True
                pri_res[ix], dua_res[ix] = pr, dr
                it_out[ix] = it
                st = check(ix, Ax, Px, Aty, pr, dr, False)
                status[ix] = st
                last_checked[ix] = True
                keep = st == UNSOLVED
            else:
                keep = np.ones(len(ix), bool)
            # adapt_rho osqp.c:488-517, auxil.c:13-74
            if s['adaptive_rho'] and ari and it % ari == 0 and keep.any():
                if not have_info:
                    Ax, Px, Aty, rp, rd, pr, dr = info(ix)
                    pri_res[ix], dua_res[ix] = pr, dr
                    it_out[ix] = it
                kk = np.nonzero(keep)[0]
                ik = ix[kk]
                with np.errstate(divide='ignore', invalid='ignore'):
                    pn = np.abs(rp[kk]).max(axis=1) / (np.maximum(np.abs(z[ik]).max(axis=1), np.abs(Ax[kk]).max(axis=1)) + DIVISION_TOL)
                    dn = np.abs(rd[kk]).max(axis=1) / (np.maximum(np.maximum(np.abs(q[ik]).max(axis=1), np.abs(Aty[kk]).max(axis=1)),
                                                                  np.abs(Px[kk]).max(axis=1)) + DIVISION_TOL)
                    rho_new = rho[ik] * np.sqrt(pn / dn)
                rho_new = np.minimum(np.maximum(rho_new, RHO_MIN), RHO_MAX)
                tol = s['adaptive_rho_tolerance']
                upd = (rho_new > rho[ik] * tol) | (rho_new < rho[ik] / tol)
                iu = ik[upd]
                if iu.size:                              # osqp_update_rho osqp.c:1268-1325
                    rho[iu] = rho_new[upd]
                    rho_vec[iu] = self._rho_vec(l[iu], u[iu], rho[iu])
                    rho_updates[iu] += 1
            active = ix[keep]
        # osqp.c:532-552: final update_info/check if the last iteration was not a check iteration
        if active.size:
            ix = active
            if not last_checked[ix].all():
                Ax, Px, Aty, rp, rd, pr, dr = info(ix)
                pri_res[ix], dua_res[ix] = pr, dr
                it_out[ix] = it
                status[ix] = check(ix, Ax, Px, Aty, pr, dr, False)
            ix = ix[status[ix] == UNSOLVED]
            if ix.size:                                  # osqp.c:563-568
                Ax, Px, Aty, rp, rd, pr, dr = info(ix)
                st = check(ix, Ax, Px, Aty, pr, dr, True)
                st[st == UNSOLVED] = MAX_ITER_REACHED
                status[ix] = st
        # compute_obj_val auxil.c:227-238 (on scaled data), store_solution :524-562, unscale_solution scaling.c:177-192
        has_sol = ~np.isin(status, [PRIMAL_INFEASIBLE, PRIMAL_INFEASIBLE_INACCURATE, DUAL_INFEASIBLE,
                                    DUAL_INFEASIBLE_INACCURATE, NON_CVX])
        obj = (0.5 * np.einsum('bi,ij,bj->b', x, P, x) + (q * x).sum(axis=1))
        if s['scaling']:
            obj = obj * cinv
        obj = np.where(np.isin(status, [PRIMAL_INFEASIBLE, PRIMAL_INFEASIBLE_INACCURATE]), OSQP_INFTY, obj)
        obj = np.where(np.isin(status, [DUAL_INFEASIBLE, DUAL_INFEASIBLE_INACCURATE]), -OSQP_INFTY, obj)
        obj = np.where(status == NON_CVX, np.nan, obj)
        xs = np.where(has_sol[:, None], D * x, np.nan)
        ys = np.where(has_sol[:, None], cinv * E * y, np.nan)
        return dict(x=xs, y=ys, obj=obj, iter=it_out, status=status.astype(np.int32), pri_res=pri_res,
                    dua_res=dua_res, rho_updates=rho_updates)


def solve_matrix_batch(P_list, A_list, q_pristine, l_pristine, u_pristine, q=None, l=None, u=None, **settings):
    """Per-instance matrices (the osqp_update_data_mat branch of the generated solve, cvxpygen/solvers/osqp.py:20-33;
    0.6.2: osqp_update_P_A, src/osqp.c:1158-1264): scale_data runs on (P_i, A_i) with the linear cost that is in the
    workspace at that moment -- the pristine one -- then the instance's own q, l, u are loaded (update_lin_cost /
    update_bounds, osqp.c:752-827) and the problem is solved from a cold start.  One AdmmOracle per instance.
    P_list / A_list: sequences of (n x n upper-triangular or symmetric) and (m x n) matrices."""
    outs = []
    for i, (P_i, A_i) in enumerate(zip(P_list, A_list)):
        orc = AdmmOracle(P_i, q_pristine, A_i, l_pristine, u_pristine, **settings)
        kw = {}
        if q is not None: kw['q'] = np.asarray(q)[i:i + 1]
        if l is not None: kw['l'] = np.asarray(l)[i:i + 1]
        if u is not None: kw['u'] = np.asarray(u)[i:i + 1]
        outs.append(orc.solve_batch(B=1, **kw) if not kw else orc.solve_batch(**kw))
    return {k: np.concatenate([np.atleast_1d(o[k]) for o in outs]) for k in outs[0]}
