"""oracle/ipm_numpy.py -- TEST INFRASTRUCTURE (oracle), not product code.

numpy restatement of the reference's SOCP hot path: ECOS 2.0.8's Mehrotra predictor-corrector interior-point method on
the homogeneous self-dual embedding with Nesterov-Todd scaling (algorithm card C.2).  Paths are
relative to cvxpygen/solvers/ecos/.

  init                      src/ecos.c:260-452     (two least-squares solves with W = I, bring2cone src/cone.c:52-95)
  computeResiduals          src/ecos.c:455-499
  updateStatistics          src/ecos.c:502-545
  checkExitConditions       src/ecos.c:179-257     (optimality branch + infeasibility certificates)
  updateScalings            src/cone.c:138-234     (LP: w = sqrt(s/z); SOC: eta, a, q)
  scale / conicProduct / conicDivision   src/cone.c:276-305, 452-513
  RHS_affine / RHS_combined src/ecos.c:648-757
  lineSearch                src/ecos.c:947-1046    (Vandenberghe's closed form for the SOC, step in [1e-6, 0.999])
  main loop                 src/ecos.c:1123-1583   (sigma = (1-alpha_aff)^3 in [1e-4, 1], step * gamma = 0.99)
  backscale                 src/ecos.c:1051-1070

Deliberate differences (documented, the optimum is unaffected): no Ruiz equilibration (src/equil.c) and a dense solve
of the exact KKT system instead of AMD + sparse LDL' with static/dynamic regularisation and iterative refinement
(src/kkt.c).  Consequently iteration counts can differ from ECOS by +-1..2; what is pinned (tests/test_socp_oracle.py)
is agreement of x, y, z with the compiled reference at its 1e-8 tolerances, i.e. well inside the 1e-5 parity bar.
"""
import numpy as np
import scipy.sparse as sp

GAMMA = 0.99
STEPMIN, STEPMAX = 1e-6, 0.999
SIGMAMIN, SIGMAMAX = 1e-4, 1.0
EPS = 1e-13


def _cones(l, q):
    out, o = [], l
    for d in q:
        out.append((o, d)); o += d
    return out


def bring2cone(r, l, q):
    """cone.c:52-95: s = r + (1 + alpha) e, alpha = largest cone violation (at least -GAMMA)."""
    a = -GAMMA
    if l:
        v = -r[:l][r[:l] <= 0]
        if v.size:
            a = max(a, v.max())
    for o, d in _cones(l, q):
        cres = r[o] - np.linalg.norm(r[o + 1:o + d])
        if cres <= 0 and -cres > a:
            a = -cres
    a += 1.0
    s = r.copy()
    s[:l] = r[:l] + a
    for o, d in _cones(l, q):
        s[o] = r[o] + a
    return s


class NT:
    """Nesterov-Todd scaling of the product cone R_+^l x Q^{q1} x ..."""

    def __init__(self, s, z, l, q):
        self.l, self.q = l, q
        self.w = np.sqrt(s[:l] / z[:l])
        self.soc = []
        for o, d in _cones(l, q):
            sk, zk = s[o:o + d], z[o:o + d]
            sres = sk[0] ** 2 - sk[1:] @ sk[1:]; zres = zk[0] ** 2 - zk[1:] @ zk[1:]
            if sres <= 0 or zres <= 0:
                raise FloatingPointError('outside cone')
            snorm, znorm = np.sqrt(sres), np.sqrt(zres)
            sb, zb = sk / snorm, zk / znorm
            eta2 = snorm / znorm
            gamma = np.sqrt(0.5 * (1.0 + sb @ zb))
            a = (sb[0] + zb[0]) / (2 * gamma)
            qv = (sb[1:] - zb[1:]) / (2 * gamma)
            self.soc.append((o, d, np.sqrt(eta2), a, qv))

    def W(self, v):                     # lambda = W z  (cone.c:276-305)
        out = np.empty_like(v)
        out[:self.l] = self.w * v[:self.l]
        for o, d, eta, a, qv in self.soc:
            zeta = qv @ v[o + 1:o + d]
            factor = v[o] + zeta / (1 + a)
            out[o] = eta * (a * v[o] + zeta)
            out[o + 1:o + d] = eta * (v[o + 1:o + d] + factor * qv)
        return out

    def W2_dense(self, m):              # W^2 as a dense matrix (eta^2 (2 wbar wbar' - J))
        M = np.zeros((m, m))
        M[np.arange(self.l), np.arange(self.l)] = self.w ** 2
        for o, d, eta, a, qv in self.soc:
            wb = np.concatenate([[a], qv])
            J = -np.eye(d); J[0, 0] = 1.0
            M[o:o + d, o:o + d] = eta ** 2 * (2 * np.outer(wb, wb) - J)
        return M


def conic_product(u, v, l, q):
    w = np.empty_like(u)
    w[:l] = u[:l] * v[:l]
    for o, d in _cones(l, q):
        w[o] = u[o:o + d] @ v[o:o + d]
        w[o + 1:o + d] = u[o] * v[o + 1:o + d] + v[o] * u[o + 1:o + d]
    return w


def conic_division(u, w, l, q):
    v = np.empty_like(u)
    v[:l] = w[:l] / u[:l]
    for o, d in _cones(l, q):
        u0, w0 = u[o], w[o]
        rho = u0 * u0 - u[o + 1:o + d] @ u[o + 1:o + d]
        zeta = u[o + 1:o + d] @ w[o + 1:o + d]
        factor = (zeta / u0 - w0) / rho
        v[o] = (u0 * w0 - zeta) / rho
        v[o + 1:o + d] = factor * u[o + 1:o + d] + w[o + 1:o + d] / u0
    return v


def line_search(lam, ds, dz, tau, dtau, kap, dkap, l, q):
    if l:
        rhomin = (ds[:l] / lam[:l]).min(); sigmamin = (dz[:l] / lam[:l]).min()
        if -sigmamin > -rhomin:
            alpha = 1.0 / (-sigmamin) if sigmamin < 0 else 1.0 / EPS
        else:
            alpha = 1.0 / (-rhomin) if rhomin < 0 else 1.0 / EPS
    else:
        alpha = 10.0
    if dtau != 0 and 0 < -tau / dtau < alpha:
        alpha = -tau / dtau
    if dkap != 0 and 0 < -kap / dkap < alpha:
        alpha = -kap / dkap
    for o, d in _cones(l, q):
        lk, dsk, dzk = lam[o:o + d], ds[o:o + d], dz[o:o + d]
        n2 = lk[0] ** 2 - lk[1:] @ lk[1:]
        if n2 <= 0:
            continue
        nrm = np.sqrt(n2); lb = lk / nrm
        step = 0.0
        for dv in (dsk, dzk):
            lt = lb[0] * dv[0] - lb[1:] @ dv[1:]
            r0 = lt / nrm
            factor = (lt + dv[0]) / (lb[0] + 1)
            r1 = (dv[1:] - factor * lb[1:]) / nrm
            step = max(step, np.linalg.norm(r1) - r0)
        if step != 0 and 1.0 / step < alpha:
            alpha = 1.0 / step
    return min(max(alpha, STEPMIN), STEPMAX)


def ecos_ipm(c, A, b, G, h, l, q, feastol=1e-8, abstol=1e-8, reltol=1e-8, maxit=100):
    """Returns dict(x, y, z, s, pcost, iter, exitflag(0 optimal, -1 maxit, 1 pinf, 2 dinf), pres, dres)."""
    A = sp.csr_matrix(A).toarray(); G = sp.csr_matrix(G).toarray()
    n, p, m = G.shape[1], A.shape[0], G.shape[0]
    D = l + len(q)

    def kkt_solve(W2, rhs):
        K = np.zeros((n + p + m, n + p + m))
        K[:n, n:n + p] = A.T; K[:n, n + p:] = G.T
        K[n:n + p, :n] = A; K[n + p:, :n] = G
        K[n + p:, n + p:] = -W2
        sol = np.linalg.solve(K + np.diag(np.r_[1e-13 * np.ones(n), -1e-13 * np.ones(p), np.zeros(m)]), rhs)
        return sol[:n], sol[n:n + p], sol[n + p:]
    # ---- init (ecos.c:260-452)
    I = np.eye(m)
    x, _, mr = kkt_solve(I, np.r_[np.zeros(n), b, h])
    s = bring2cone(-mr, l, q)
    _, y, zb = kkt_solve(I, np.r_[-c, np.zeros(p), np.zeros(m)])
    z = bring2cone(zb, l, q)
    kap = tau = 1.0
    resx0, resy0, resz0 = max(1, np.linalg.norm(c)), max(1, np.linalg.norm(b)), max(1, np.linalg.norm(h))
    exitflag, it = -1, 0
    pres = dres = np.nan
    for it in range(maxit + 1):
        rx = -A.T @ y - G.T @ z - c * tau
        ry = A @ x - b * tau
        rz = s + G @ x - h * tau
        cx, by, hz = c @ x, b @ y, h @ z
        rt = kap + cx + by + hz
        nx, ny, ns, nz = (np.linalg.norm(v) for v in (x, y, s, z))
        gap = s @ z
        mu = (gap + kap * tau) / (D + 1)
        pcost, dcost = cx / tau, -(hz + by) / tau
        relgap = gap / (-pcost) if pcost < 0 else (gap / dcost if dcost > 0 else np.nan)
        pres = max(np.linalg.norm(ry) / max(resy0 + nx, 1), np.linalg.norm(rz) / max(resz0 + nx + ns, 1)) / tau
        dres = np.linalg.norm(rx) / max(resx0 + ny + nz, 1) / tau
        hresx = np.linalg.norm(-A.T @ y - G.T @ z); hresy = np.linalg.norm(A @ x); hresz = np.linalg.norm(s + G @ x)
        pinfres = hresx / max(ny + nz, 1) if (hz + by) / max(ny + nz, 1) < -reltol else np.nan
        dinfres = max(hresy / max(nx, 1), hresz / max(nx + ns, 1)) if cx / max(nx, 1) < -reltol else np.nan
        if (-cx > 0 or -by - hz >= -abstol) and pres < feastol and dres < feastol and (gap < abstol or relgap < reltol):
            exitflag = 0; break
        if not np.isnan(dinfres) and dinfres < feastol and tau < kap:
            exitflag = 2; break
        if (not np.isnan(pinfres) and pinfres < feastol and tau < kap) or (tau < feastol and kap < feastol and pinfres < feastol):
            exitflag = 1; break
        if it == maxit:
            exitflag = -1; break
        nt = NT(s, z, l, q)
        lam = nt.W(z)
        W2 = nt.W2_dense(m)
        x1, y1, z1 = kkt_solve(W2, np.r_[-c, b, h])
        x2, y2, z2 = kkt_solve(W2, np.r_[rx, -ry, s - rz])
        dtau_denom = kap / tau - c @ x1 - b @ y1 - h @ z1
        dtauaff = (rt - kap + c @ x2 + b @ y2 + h @ z2) / dtau_denom
        dzaff = z2 + dtauaff * z1
        Wdz = nt.W(dzaff)
        dsW = -Wdz - lam
        dkapaff = -kap - kap / tau * dtauaff
        a_aff = line_search(lam, dsW, Wdz, tau, dtauaff, kap, dkapaff, l, q)
        sigma = min(max((1 - a_aff) ** 3, SIGMAMIN), SIGMAMAX)
        # combined direction (ecos.c:688-757)
        ds1 = conic_product(lam, lam, l, q) + conic_product(dsW, Wdz, l, q)
        e = np.zeros(m); e[:l] = 1.0
        for o, d in _cones(l, q):
            e[o] = 1.0
        ds1 -= sigma * mu * e
        lam_div = conic_division(lam, ds1, l, q)
        rhs_z = -(1 - sigma) * rz + nt.W(lam_div)
        x2, y2, z2 = kkt_solve(W2, np.r_[(1 - sigma) * rx, -(1 - sigma) * ry, rhs_z])
        bkap = kap * tau + dkapaff * dtauaff - sigma * mu
        dtau = ((1 - sigma) * rt - bkap / tau + c @ x2 + b @ y2 + h @ z2) / dtau_denom
        dx, dy, dz = x2 + dtau * x1, y2 + dtau * y1, z2 + dtau * z1
        Wdz = nt.W(dz)
        dsW = -(lam_div + Wdz)
        dkap = -(bkap + kap * dtau) / tau
        step = line_search(lam, dsW, Wdz, tau, dtau, kap, dkap, l, q) * GAMMA
        ds = nt.W(dsW)
        x += step * dx; y += step * dy; z += step * dz; s += step * ds
        kap += step * dkap; tau += step * dtau
    return dict(x=x / tau, y=y / tau, z=z / tau, s=s / tau, pcost=c @ x / tau, iter=it, exitflag=exitflag, pres=pres, dres=dres)


# ======================================================================================================================
# Exact restatement: what the compiled reference does, step for step (equilibration, stretched KKT with static and
# dynamic regularisation, iterative refinement with its stopping rules, best-iterate safeguards, back-scaling).
# Differences from the C code are limited to floating-point summation order and the elimination ordering of the LDL'
# (any ordering solves the same regularised system; the refinement loop removes the ordering-dependent error).
# ======================================================================================================================
DELTASTAT, DELTA, EPS_DYN = 7e-8, 2e-7, 1e-13
NITREF, IRERRFACT, LINSYSACC = 9, 6, 1e-14
SAFEGUARD = 500
FTOL_INACC, ATOL_INACC, RTOL_INACC = 1e-4, 5e-5, 5e-5
NOT_CONVERGED = -87


def ruiz_equilibrate(A, G, l, q, iters=3):
    """use_ruiz_equilibration, src/equil.c:210-340: `iters` passes of sqrt(max-abs) row / column scaling; the rows of
    one second-order cone share the SUM of their row maxima.  Returns (A_eq, G_eq, xequil, Aequil, Gequil)."""
    A = sp.csc_matrix(A, dtype=float, copy=True); G = sp.csc_matrix(G, dtype=float, copy=True)
    A.sort_indices(); G.sort_indices()
    n, p, m = G.shape[1], A.shape[0], G.shape[0]
    xe, Ae, Ge = np.ones(n), np.ones(p), np.ones(m)
    Acol = np.repeat(np.arange(n), np.diff(A.indptr)); Gcol = np.repeat(np.arange(n), np.diff(G.indptr))
    for _ in range(iters):
        xt, At, Gt = np.zeros(n), np.zeros(p), np.zeros(m)
        np.maximum.at(xt, Acol, np.abs(A.data)); np.maximum.at(xt, Gcol, np.abs(G.data))
        np.maximum.at(At, A.indices, np.abs(A.data)); np.maximum.at(Gt, G.indices, np.abs(G.data))
        for o, d in _cones(l, q):
            tot = 0.0
            for j in range(d):
                tot += Gt[o + j]
            Gt[o:o + d] = tot
        xt = np.where(np.abs(xt) < 1e-6, 1.0, np.sqrt(xt))
        At = np.where(np.abs(At) < 1e-6, 1.0, np.sqrt(At))
        Gt = np.where(np.abs(Gt) < 1e-6, 1.0, np.sqrt(Gt))
        A.data = A.data / At[A.indices]; G.data = G.data / Gt[G.indices]          # rows first ...
        A.data = A.data / xt[Acol]; G.data = G.data / xt[Gcol]                    # ... then columns
        xe *= xt; Ae *= At; Ge *= Gt
    return A, G, xe, Ae, Ge


def _safediv(x, y):
    return x / np.where(y < EPS, EPS, y) if isinstance(y, np.ndarray) else (x / EPS if y < EPS else x / y)


class _Stretch:
    """Index bookkeeping of the 'stretched' KKT system (CONEMODE 0): every second-order cone of size d occupies
    d + 2 rows, the two extra ones carrying the sparse representation of its NT scaling (src/preproc.c:77-330)."""

    def __init__(self, n, p, l, q):
        self.n, self.p, self.l, self.q = n, p, l, list(q)
        m = l + sum(q)
        self.m, self.mt = m, m + 2 * len(q)
        self.nK = n + p + self.mt
        zmap = np.zeros(m, dtype=int); zmap[:l] = np.arange(l)
        self.blocks = []
        o, so = l, l
        for d in q:
            zmap[o:o + d] = so + np.arange(d)
            self.blocks.append((o, so, d)); o += d; so += d + 2
        self.zmap = zmap
        sign = np.r_[np.ones(n), -np.ones(p), -np.ones(self.mt)]
        for o, so, d in self.blocks:
            sign[n + p + so + d + 1] = 1.0
        self.sign = sign

    def stretch(self, vz):
        out = np.zeros(self.mt); out[self.zmap] = vz
        return out


class _Scaling:
    """updateScalings, src/cone.c:138-234."""

    def __init__(self, s, z, l, q):
        self.l, self.q = l, q
        self.v = _safediv(s[:l], z[:l]); self.w = np.sqrt(self.v)
        self.soc = []
        self.ok = True
        for o, d in _cones(l, q):
            sk, zk = s[o:o + d], z[o:o + d]
            sres = sk[0] * sk[0] - sk[1:] @ sk[1:]; zres = zk[0] * zk[0] - zk[1:] @ zk[1:]
            if sres <= 0 or zres <= 0:
                self.ok = False; return
            snorm, znorm = np.sqrt(sres), np.sqrt(zres)
            skbar, zkbar = _safediv(sk, snorm), _safediv(zk, znorm)
            eta2 = _safediv(snorm, znorm); eta = np.sqrt(eta2)
            gamma = np.sqrt(0.5 * (1.0 + skbar @ zkbar))
            o2g = _safediv(0.5, gamma)
            a = o2g * (skbar[0] + zkbar[0])
            qv = o2g * (skbar[1:] - zkbar[1:])
            w = qv @ qv
            temp = 1.0 + a
            c = 1.0 + a + _safediv(w, temp)
            dd = 1 + _safediv(2, temp) + _safediv(w, temp * temp)
            d1 = max(0.0, 0.5 * (a * a + w * (1.0 - _safediv(c * c, 1.0 + w * dd))))
            u0sq = a * a + w - d1
            u0 = np.sqrt(u0sq)
            c2byu02 = _safediv(c * c, u0sq)
            if c2byu02 - dd <= 0:
                self.ok = False; return
            v1 = np.sqrt(c2byu02 - dd); u1 = np.sqrt(c2byu02)
            self.soc.append(dict(o=o, d=d, eta2=eta2, eta=eta, a=a, q=qv, w=w, d1=d1, u0=u0, u1=u1, v1=v1))

    def scale(self, z):                                    # lambda = W z, src/cone.c:276-305
        out = np.empty_like(z)
        out[:self.l] = self.w * z[:self.l]
        for c in self.soc:
            o, d, qv, a, eta = c['o'], c['d'], c['q'], c['a'], c['eta']
            zeta = qv @ z[o + 1:o + d]
            factor = z[o] + _safediv(zeta, 1 + a)
            out[o] = eta * (a * z[o] + zeta)
            out[o + 1:o + d] = eta * (z[o + 1:o + d] + factor * qv)
        return out


def _kkt_matrix(st, A, G, sc):
    """Dense stretched KKT matrix: kkt_init (src/kkt.c:373-445) when sc is None, kkt_update (:271-357) otherwise, on
    the skeleton of createKKT_U (src/preproc.c:77-330): +delta on the x block, -delta on the y block."""
    n, p, l = st.n, st.p, st.l
    K = np.zeros((st.nK, st.nK))
    K[np.arange(n), np.arange(n)] = DELTASTAT
    K[n + np.arange(p), n + np.arange(p)] = -DELTASTAT
    K[:n, n:n + p] = A.T; K[n:n + p, :n] = A
    zr = n + p + st.zmap
    K[zr, :n] = G; K[:n, zr] = G.T
    zz = n + p
    if sc is None:
        K[zz + np.arange(l), zz + np.arange(l)] = -1.0
        for o, so, d in st.blocks:
            i = zz + so + np.arange(d)
            K[i, i] = -1.0
            K[zz + so + d, zz + so + d] = -1.0; K[zz + so + d + 1, zz + so + d + 1] = 1.0
        return K
    K[zz + np.arange(l), zz + np.arange(l)] = -sc.v - DELTASTAT
    for (o, so, d), c in zip(st.blocks, sc.soc):
        e2, qv = c['eta2'], c['q']
        i = zz + so + np.arange(d)
        K[i, i] = -e2 - DELTASTAT
        K[i[0], i[0]] = -e2 * c['d1'] - DELTASTAT
        iv, iu = zz + so + d, zz + so + d + 1
        K[i[1:], iv] = -e2 * c['v1'] * qv; K[iv, i[1:]] = K[i[1:], iv]
        K[iv, iv] = -e2
        K[i[0], iu] = -e2 * c['u0']; K[iu, i[0]] = K[i[0], iu]
        K[i[1:], iu] = -e2 * c['u1'] * qv; K[iu, i[1:]] = K[i[1:], iu]
        K[iu, iu] = e2 + DELTASTAT
    return K


def _ldl(Kp, pattern, sign):
    """LDL_numeric2 (external/ldl/src/ldl.c:266-360) with its dynamic regularisation, on the permuted dense matrix."""
    nK = Kp.shape[0]
    S = Kp.copy(); L = np.zeros_like(S); D = np.zeros(nK)
    for k in range(nK):
        dk = S[k, k]
        if sign[k] * dk <= EPS_DYN:
            dk = sign[k] * DELTA
        D[k] = dk
        rows = pattern[k]
        if rows.size:
            col = S[rows, k]
            lk = col / dk
            L[rows, k] = lk
            S[np.ix_(rows, rows)] -= np.outer(lk, col)
    return L, D


class EcosExact:
    """ECOS_setup + per-instance ECOS_updateData/ECOS_solve for a family with fixed A, G (src/ecos.c, src/preproc.c)."""

    def __init__(self, A, G, l, q, order=None, feastol=1e-8, abstol=1e-8, reltol=1e-8, maxit=100):
        self.l, self.q = l, list(q)
        self.Aeq, self.Geq, self.xe, self.Ae, self.Ge = ruiz_equilibrate(A, G, l, q)
        self.A, self.G = self.Aeq.toarray(), self.Geq.toarray()
        self.n, self.p, self.m = self.G.shape[1], self.A.shape[0], self.G.shape[0]
        self.st = _Stretch(self.n, self.p, l, q)
        self.tol = (feastol, abstol, reltol); self.maxit = maxit
        self.D = l + len(q)
        K0 = _kkt_matrix(self.st, self.A, self.G, None)
        if order is None:
            from cvxpygen_b200.offline import kkt as _k            # fill-reducing ordering only
            pat = sp.csr_matrix((K0 != 0).astype(float))
            # the v / u columns are structurally non-zero although kkt_init writes zeros there
            pat = pat.tolil()
            zz = self.n + self.p
            for o, so, d in self.st.blocks:
                for r in range(1, d):
                    pat[zz + so + r, zz + so + d] = 1; pat[zz + so + d, zz + so + r] = 1
                for r in range(d):
                    pat[zz + so + r, zz + so + d + 1] = 1; pat[zz + so + d + 1, zz + so + r] = 1
            pat = sp.csr_matrix(pat)
            order = _k.minimum_degree_order(pat)
            struct, _ = _k._symbolic(pat, order)
            self.pattern = struct
        self.perm = np.asarray(order)
        self.nitref_log = []

    # ---- linear algebra
    def _factor(self, sc):
        K = _kkt_matrix(self.st, self.A, self.G, sc)
        pm = self.perm
        self.L, self.Dg = _ldl(K[np.ix_(pm, pm)], self.pattern, self.st.sign[pm])
        self.Lu = np.tril(self.L, -1) + np.eye(self.st.nK)

    def _ldl_solve(self, b):
        from scipy.linalg import solve_triangular
        pm = self.perm
        y = solve_triangular(self.Lu, b[pm], lower=True, unit_diagonal=True)
        y = y / self.Dg
        xp = solve_triangular(self.Lu.T, y, lower=False, unit_diagonal=True)
        out = np.empty_like(b); out[pm] = xp
        return out

    def _kkt_solve(self, b, sc, isinit):
        """kkt_solve, src/kkt.c:87-265.  b and the returned vector live in the natural stretched ordering."""
        st, n, p, l = self.st, self.n, self.p, self.l
        zz = n + p
        bnorm = 1.0 + np.abs(b).max()
        thr = bnorm * LINSYSACC
        Px = self._ldl_solve(b)
        nerr_prev = np.nan
        dPx = None
        k = 0
        while True:
            dx, dy, tz = Px[:n], Px[n:zz], Px[zz:]
            dz = tz[st.zmap]
            ex = b[:n] - DELTASTAT * dx - self.A.T @ dy - self.G.T @ dz
            ey = b[n:zz] + DELTASTAT * dy - self.A @ dx
            Gdx = self.G @ dx
            ez = np.zeros(st.mt)
            sgn = np.ones(self.m)
            for o, so, d in st.blocks:
                sgn[o + d - 1] = -1.0
            ez[st.zmap] = b[zz + st.zmap] - Gdx + sgn * DELTASTAT * dz
            if isinit:
                ez += tz
            else:
                ez[:l] += sc.v * tz[:l]
                for (o, so, d), c in zip(st.blocks, sc.soc):
                    e2, qv = c['eta2'], c['q']
                    x1, x2, x3, x4 = tz[so], tz[so + 1:so + d], tz[so + d], tz[so + d + 1]
                    qtx2 = qv @ x2
                    ez[so] += e2 * (c['d1'] * x1 + c['u0'] * x4)
                    ez[so + 1:so + d] += e2 * (x2 + (c['v1'] * x3 + c['u1'] * x4) * qv)
                    ez[so + d] += e2 * (c['v1'] * qtx2 + x3)
                    ez[so + d + 1] += e2 * (c['u0'] * x1 + c['u1'] * qtx2 - x4)
            nerr = max(np.abs(ex).max(), np.abs(ez).max(), np.abs(ey).max() if p else 0.0)
            if k > 0 and nerr > nerr_prev:
                Px = Px - dPx; k -= 1
                break
            if k == NITREF or nerr < thr or (k > 0 and nerr_prev < IRERRFACT * nerr):
                break
            nerr_prev = nerr
            dPx = self._ldl_solve(np.r_[ex, ey, ez])
            Px = Px + dPx
            k += 1
        self.nitref_log.append(k)
        return Px[:n].copy(), Px[n:zz].copy(), Px[zz:][st.zmap].copy()

    # ---- solver
    def _stats(self, W):
        A, G = self.A, self.G
        x, y, z, s, tau, kap, c, b, h = (W[k] for k in ('x', 'y', 'z', 's', 'tau', 'kap', 'c', 'b', 'h'))
        hrx = -A.T @ y - G.T @ z; W['hresx'] = np.linalg.norm(hrx); W['rx'] = hrx - tau * c
        hry = A @ x; W['hresy'] = np.linalg.norm(hry); W['ry'] = hry - tau * b
        hrz = s + G @ x; W['hresz'] = np.linalg.norm(hrz); W['rz'] = hrz - tau * h
        W['cx'], W['by'], W['hz'] = c @ x, b @ y, h @ z
        W['rt'] = kap + W['cx'] + W['by'] + W['hz']
        nx, ny, ns, nz = (np.linalg.norm(v) for v in (x, y, s, z))
        I = {}
        I['gap'] = s @ z
        I['mu'] = (I['gap'] + kap * tau) / (self.D + 1)
        I['kapovert'] = kap / tau
        I['pcost'] = W['cx'] / tau; I['dcost'] = -(W['hz'] + W['by']) / tau
        I['relgap'] = I['gap'] / (-I['pcost']) if I['pcost'] < 0 else (I['gap'] / I['dcost'] if I['dcost'] > 0 else np.nan)
        nry = np.linalg.norm(W['ry']) / max(W['resy0'] + nx, 1) if self.p else 0.0
        nrz = np.linalg.norm(W['rz']) / max(W['resz0'] + nx + ns, 1)
        I['pres'] = max(nry, nrz) / tau
        I['dres'] = np.linalg.norm(W['rx']) / max(W['resx0'] + ny + nz, 1) / tau
        reltol = self.tol[2]
        I['pinfres'] = W['hresx'] / max(ny + nz, 1) if (W['hz'] + W['by']) / max(ny + nz, 1) < -reltol else np.nan
        I['dinfres'] = max(W['hresy'] / max(nx, 1), W['hresz'] / max(nx + ns, 1)) if W['cx'] / max(nx, 1) < -reltol else np.nan
        W['info'] = I

    def _exit(self, W, mode):
        feastol, abstol, reltol = self.tol if mode == 0 else (FTOL_INACC, ATOL_INACC, RTOL_INACC)
        I = W['info']
        if (-W['cx'] > 0 or -W['by'] - W['hz'] >= -abstol) and (I['pres'] < feastol and I['dres'] < feastol) and \
                (I['gap'] < abstol or I['relgap'] < reltol):
            return 0 + mode
        if I['dinfres'] < feastol and W['tau'] < W['kap']:
            return 2 + mode
        if (I['pinfres'] < feastol and W['tau'] < W['kap']) or \
                (W['tau'] < self.tol[0] and W['kap'] < self.tol[0] and I['pinfres'] < self.tol[0]):
            return 1 + mode
        return NOT_CONVERGED

    @staticmethod
    def _better(a, b):
        """compareStatistics, src/ecos.c:61-98 (`x != ECOS_NAN` is always true in C)."""
        g = a['gap'] > 0 and b['gap'] > 0 and a['gap'] < b['gap']
        mu = a['mu'] > 0 and a['mu'] < b['mu']
        if a['kapovert'] > 1:
            return g and (a['pinfres'] > 0 and a['pinfres'] < b['pres']) and mu
        return g and (a['pres'] > 0 and a['pres'] < b['pres']) and (a['dres'] > 0 and a['dres'] < b['dres']) and \
            (a['kapovert'] > 0 and a['kapovert'] < b['kapovert']) and mu

    def solve(self, c, b, h):
        l, q, n, p, m, st = self.l, self.q, self.n, self.p, self.m, self.st
        W = dict(c=np.asarray(c, float) / self.xe, b=np.asarray(b, float) / self.Ae, h=np.asarray(h, float) / self.Ge)
        c, b, h = W['c'], W['b'], W['h']
        self.nitref_log = []
        # ---- init, src/ecos.c:260-452
        self._factor(None)
        rhs1 = np.r_[np.zeros(n), b, st.stretch(h)]
        x, _, mr = self._kkt_solve(rhs1, None, True)
        s = bring2cone(-mr, l, q)
        _, y, zb = self._kkt_solve(np.r_[-c, np.zeros(p), np.zeros(st.mt)], None, True)
        z = bring2cone(zb, l, q)
        rhs1[:n] = -c
        W.update(x=x, y=y, z=z, s=s, tau=1.0, kap=1.0, resx0=max(1, np.linalg.norm(c)), resy0=max(1, np.linalg.norm(b)),
                 resz0=max(1, np.linalg.norm(h)))
        best = None
        pres_prev = np.nan
        step = 0.0
        exitcode = -7

        def restore():
            for k in ('x', 'y', 'z', 's', 'tau', 'kap', 'cx', 'by', 'hz'):
                W[k] = best[k] if np.isscalar(best[k]) else best[k].copy()
            it = W['info'].get('iter')
            W['info'] = dict(best['info'])

        def save():
            nonlocal best
            best = {k: (W[k] if np.isscalar(W[k]) else W[k].copy()) for k in ('x', 'y', 'z', 's', 'tau', 'kap', 'cx', 'by', 'hz')}
            best['info'] = dict(W['info'])

        it = 0
        while True:
            self._stats(W)
            I = W['info']
            if it > 0 and (I['pres'] > SAFEGUARD * pres_prev or I['gap'] < 0):
                restore()
                exitcode = self._exit(W, 10)
                if exitcode == NOT_CONVERGED:
                    exitcode = -2
                break
            pres_prev = I['pres']
            exitcode = self._exit(W, 0)
            if exitcode != NOT_CONVERGED:
                break
            if it > 0 and step == STEPMIN * GAMMA:
                restore(); exitcode = self._exit(W, 10)
                if exitcode == NOT_CONVERGED:
                    exitcode = -2
                break
            if it == self.maxit:
                if not self._better(I, best['info']):
                    restore()
                exitcode = self._exit(W, 10)
                if exitcode == NOT_CONVERGED:
                    exitcode = -1
                break
            if np.isnan(I['pcost']):
                if not self._better(I, best['info']):
                    restore()
                exitcode = self._exit(W, 10)
                if exitcode == NOT_CONVERGED:
                    exitcode = -2
                break
            if it == 0 or self._better(I, best['info']):
                save()
            x, y, z, s, tau, kap = (W[k] for k in ('x', 'y', 'z', 's', 'tau', 'kap'))
            sc = _Scaling(s, z, l, q)
            if not sc.ok:
                restore(); exitcode = self._exit(W, 10)
                if exitcode == NOT_CONVERGED:
                    exitcode = -3
                break
            lam = sc.scale(z)
            self._factor(sc)
            x1, y1, z1 = self._kkt_solve(rhs1, sc, False)
            rx, ry, rz, rt = W['rx'], W['ry'], W['rz'], W['rt']
            rhs2 = np.r_[rx, -ry, st.stretch(s - rz)]
            x2, y2, z2 = self._kkt_solve(rhs2, sc, False)
            dtau_denom = kap / tau - c @ x1 - b @ y1 - h @ z1
            dtauaff = (rt - kap + c @ x2 + b @ y2 + h @ z2) / dtau_denom
            z2 = z2 + dtauaff * z1
            Wdz = sc.scale(z2)
            dsW = -Wdz - lam
            dkapaff = -kap - kap / tau * dtauaff
            step_aff = line_search(lam, dsW, Wdz, tau, dtauaff, kap, dkapaff, l, q)
            sigma = min(max((1.0 - step_aff) ** 3, SIGMAMIN), SIGMAMAX)
            mu = I['mu']
            ds1 = conic_product(lam, lam, l, q) + conic_product(dsW, Wdz, l, q)
            ds1[:l] -= sigma * mu
            for o, d in _cones(l, q):
                ds1[o] -= sigma * mu
            dsW = conic_division(lam, ds1, l, q)
            ds1 = sc.scale(dsW)
            oms = 1.0 - sigma
            rhs2 = np.r_[oms * rx, oms * (-ry), st.stretch(-oms * rz + ds1)]
            x2, y2, z2 = self._kkt_solve(rhs2, sc, False)
            bkap = kap * tau + dkapaff * dtauaff - sigma * mu
            dtau = (oms * rt - bkap / tau + c @ x2 + b @ y2 + h @ z2) / dtau_denom
            x2 = x2 + dtau * x1; y2 = y2 + dtau * y1; z2 = z2 + dtau * z1
            Wdz = sc.scale(z2)
            dsW = -(dsW + Wdz)
            dkap = -(bkap + kap * dtau) / tau
            step = line_search(lam, dsW, Wdz, tau, dtau, kap, dkap, l, q) * GAMMA
            ds = sc.scale(dsW)
            W['x'] = x + step * x2; W['y'] = y + step * y2; W['z'] = z + step * z2; W['s'] = s + step * ds
            W['kap'] = kap + step * dkap; W['tau'] = tau + step * dtau
            it += 1
        tau = W['tau']
        I = W['info']
        return dict(x=W['x'] / (self.xe * tau), y=W['y'] / (self.Ae * tau), z=W['z'] / (self.Ge * tau),
                    s=W['s'] * (self.Ge / tau), pcost=I['pcost'], iter=it, exitflag=exitcode, pres=I['pres'],
                    dres=I['dres'], nitref=list(self.nitref_log))
