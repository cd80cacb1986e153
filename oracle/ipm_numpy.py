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
