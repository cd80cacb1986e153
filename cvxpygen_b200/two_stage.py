"""Two-stage gradient for the conic backend (VERDICT r1 missing item 3).

The reference differentiates through a solve by a conic solver like this (`gradient=True` with solver != OSQP:
cvxpygen/generator.py:86-88, cvxpygen/canonicalizer.py:54-65 `canonicalize_two_stage`, :334-406 `_get_osqp_problem`, `_merge`):

  stage 1  canonicalise the user's problem into the OSQP QP form (it must be a QP whose P is diagonal and parameter-free:
           "extended DPP", canonicalizer.py:338-345);
  stage 2  re-state that QP as  minimise 1/2 |sqrt(P_diag) x|^2 + q'x  s.t.  l <= A[:n_eq] x,  A x <= u  and canonicalise IT for
           the conic solver; the user parameters reach the conic data through the product of the two affine maps (`_merge`);
  solve    with the conic solver; bring the solution back to the QP form -- x = the `osqp_x` block of the conic primal,
           y_i = z_u,i - z_l,i on the two-sided rows, z_u,i on the others (cvxpygen/writer.py:177-206,
           cpg_retrieve_intermediate_primal / _dual);
  backward the QP backward pass (cpg_osqp_gradient, templates/cpg_osqp_grad_compute.c.jinja2:432-531) at that (x, y), then the
           un-canonicalisation through the FIRST stage's maps.

Here: `conic_family_of_qp` is stage 2 (hand-derived ECOS form, cvxpy being absent: variables [x ; s], objective
1/2 s + q'x, one second-order cone  (s + 1, 2 sqrt(P_ii) x_i ..., s - 1)  i.e. |sqrt(P) x|^2 <= s, LP rows [-A_lower ; A_upper] x
<= [-l ; u] over the rows bounded below / above); `generate_two_stage` emits the IPM-CUDA library of that conic family plus the ADMM-CUDA library of the QP family
(whose backward kernel `qp_grad_kernel` is the second stage), `TwoStageModule` chains them batch-wise on the GPU.
"""
import os
from types import SimpleNamespace

import numpy as np
import scipy.sparse as sp

from .ir import CanonFamily, UserParam, UserVar, UserDual

OSQP_INFTY = 1e30


def conic_family_of_qp(fam: CanonFamily, batch_params=None, name=None) -> CanonFamily:
    """Stage 2: the ECOS-form restatement of a QP family with diagonal, parameter-free P (canonicalizer.py:334-362)."""
    if fam.solver_type != 'quadratic':
        raise ValueError('two-stage gradients start from the QP canonical form')
    if fam.changes('P'):
        raise ValueError('Problem does not follow extended DPP rules for differentiation with general solvers '
                         '(other than OSQP). Quadratics cannot be multiplied with parameters.')      # canonicalizer.py:339-343
    P = fam.canon_matrix('P')
    r, c = P.nonzero()
    assert np.all(r == c), 'P must be diagonal'                                                     # canonicalizer.py:345-346
    if fam.changes('A', batch_params):
        raise ValueError('per-instance matrices are not generated for the conic backend yet')
    if batch_params is not None:           # the other parameters are shared by the batch: constants of this generated code
        fam.param_columns(batch_params)
    n, n_eq, m = fam.n_var, fam.n_eq, fam.n_eq + fam.n_ineq
    A = fam.canon_matrix('A').tocsr()
    pd = np.asarray(P.diagonal())
    nzp = np.nonzero(pd)[0]
    u0 = np.clip(fam.canon_data('u'), -OSQP_INFTY, OSQP_INFTY)
    l0 = np.clip(fam.canon_data('l'), -OSQP_INFTY, OSQP_INFTY)
    # rows bounded above / below.  cvxpy's QP form has l = -inf on every inequality row, so the reference takes lower bounds from
    # the first n_eq rows only (canonicalizer.py:351,359); hand-built families may carry two-sided inequality rows: kept as well
    upp = lambda M, v0, sgn: np.nonzero((sgn * v0 < OSQP_INFTY * 1e-4) | np.asarray(np.diff(sp.csr_matrix(M).indptr) > 0))[0]
    keep_u = upp(fam.maps['u'][:, :-1], u0, 1.0)
    keep_l = upp(fam.maps['l'][:, :-1], l0, -1.0)
    assert np.all(np.isin(np.arange(n_eq), keep_l)), 'equality rows are bounded below'
    n_lp = len(keep_l) + len(keep_u)
    q_soc = len(nzp) + 2
    nv = n + 1
    # G v <=_K h:  LP rows [-A_eq ; A_keep] x ;  cone rows  (h - G v) = (1 + s, 2 sqrt(P_ii) x_i, -1 + s)
    G_lp = sp.vstack([-A[keep_l], A[keep_u]]) if n_lp else sp.csr_matrix((0, n))
    G_lp = sp.hstack([G_lp, sp.csr_matrix((n_lp, 1))])
    rows = [sp.csr_matrix(([-1.0], ([0], [n])), shape=(1, nv)),
            sp.csr_matrix((-2.0 * np.sqrt(pd[nzp]), (np.arange(len(nzp)), nzp)), shape=(len(nzp), nv)),
            sp.csr_matrix(([-1.0], ([0], [n])), shape=(1, nv))]
    G = sp.vstack([G_lp] + rows).tocsc(); G.sort_indices()
    mc = G.shape[0]
    n_theta = fam.n_theta
    # affine maps = second-stage (identity-like) maps times the first-stage maps (canonicalizer.py:376-383)
    Mq, Ml, Mu = sp.csr_matrix(fam.maps['q']), sp.csr_matrix(fam.maps['l']), sp.csr_matrix(fam.maps['u'])
    const_row = lambda v: sp.csr_matrix(([v], ([0], [n_theta - 1])), shape=(1, n_theta)) if v else sp.csr_matrix((1, n_theta))
    map_c = sp.vstack([Mq, const_row(0.5)]).tocsr()
    map_h = sp.vstack([-Ml[keep_l], Mu[keep_u], const_row(1.0), sp.csr_matrix((len(nzp), n_theta)), const_row(-1.0)]).tocsr()
    map_G = sp.csr_matrix((G.data, (np.arange(G.nnz), np.full(G.nnz, n_theta - 1))), shape=(G.nnz, n_theta))
    maps = {'c': map_c, 'h': map_h, 'G': map_G, 'b': sp.csr_matrix((0, n_theta)), 'A': sp.csr_matrix((0, n_theta)),
            'd': sp.csr_matrix(fam.maps['d']) if 'd' in fam.maps else sp.csr_matrix((1, n_theta))}
    Ac = sp.csc_matrix((0, nv))
    pat = lambda M: (M.indices.astype(np.int32), M.indptr.astype(np.int32), M.shape)
    # user-level variables: the QP family's, inside the x block (offset 0); duals of the conic form: z of the LP rows
    variables = [UserVar(v.name, v.shape, np.asarray(v.indices)) for v in fam.variables]
    duals = [UserDual('z_l', 'z', (len(keep_l),), np.arange(len(keep_l))),
             UserDual('z_u', 'z', (len(keep_u),), len(keep_l) + np.arange(len(keep_u)))]
    params = [UserParam(p.name, p.shape, p.size, p.col, np.array(p.default, dtype=float)) for p in fam.params]
    out = CanonFamily(name or fam.name + '_conic', 'conic', nv, 0, mc, params, maps, {'A': pat(Ac), 'G': pat(G)}, variables,
                      [d for d in duals if len(d.indices)], is_maximization=fam.is_maximization,
                      cone_dims={'l': int(n_lp), 'q': [int(q_soc)]})
    out.two_stage = dict(keep_l=keep_l, keep_u=keep_u, n=n, m=m)
    return out


def intermediate_solution(cfam: CanonFamily, x_conic, z_conic):
    """Conic solution -> QP-form (x, y): cpg_retrieve_intermediate_primal / _dual (cvxpygen/writer.py:177-206)."""
    ts = cfam.two_stage
    n, m, kl, ku = ts['n'], ts['m'], ts['keep_l'], ts['keep_u']
    x = np.ascontiguousarray(x_conic[:, :n])
    y = np.zeros((x_conic.shape[0], m))
    y[:, ku] = z_conic[:, len(kl):len(kl) + len(ku)]
    y[:, kl] -= z_conic[:, :len(kl)]
    return x, y


def generate_two_stage(fam_qp: CanonFamily, code_dir: str, batch_params=None, prefix='', verbose=False, compile=True):
    """`generate_code(problem, solver='IPM-CUDA', gradient=True)`: <code_dir> = the IPM-CUDA library of the conic restatement,
    <code_dir>/gradient = the ADMM-CUDA library of the QP form (role of c/osqp_code + the `gradient_` prefix in the reference's
    layout, cvxpygen/generator.py:131-139).  Returns (conic family, conic setup, QP setup)."""
    from . import codegen, codegen_ipm
    from .offline.qp_setup import setup_qp_family
    from .offline.socp_setup import setup_socp_family, DEFAULT_THREADS
    import pickle
    cfam = conic_family_of_qp(fam_qp, batch_params)
    csetup = setup_socp_family(cfam, batch_params, threads=DEFAULT_THREADS)
    codegen_ipm.write_ipm_code(csetup, code_dir, prefix=prefix)
    qsetup = setup_qp_family(fam_qp, batch_params)
    gdir = os.path.join(code_dir, 'gradient')
    codegen.write_code(qsetup, gdir, prefix=('gradient_' + prefix) if prefix else 'gradient')
    with open(os.path.join(code_dir, 'cpg_two_stage.pkl'), 'wb') as f:
        pickle.dump(dict(two_stage=cfam.two_stage), f)
    if compile:
        codegen_ipm.compile_ipm_code(code_dir, verbose=verbose)
        codegen.compile_code(gdir, verbose=verbose)
    return cfam, csetup, qsetup


class TwoStageModule:
    """Forward through the conic library, backward through the QP library's backward kernel, both batched on one device."""

    def __init__(self, code_dir, device=0):
        import pickle
        from . import runtime
        self.conic = runtime.load(code_dir, device)
        self.qp = runtime.load(os.path.join(code_dir, 'gradient'), device)
        with open(os.path.join(code_dir, 'cpg_two_stage.pkl'), 'rb') as f:
            self.ts = pickle.load(f)['two_stage']

    def solve_batch(self, params, **settings):
        """-> result of the conic solve + QP-form solution `sol_x`, `sol_y` (what the backward pass differentiates at)."""
        r = self.conic.solve_batch(params, return_canonical=True, **settings)
        fake = SimpleNamespace(two_stage=self.ts)
        x, y = intermediate_solution(fake, r.sol_x, r.sol_z)
        return SimpleNamespace(cpg_prim=r.cpg_prim, cpg_info=r.cpg_info, prim=r.prim, conic=r, sol_x=x, sol_y=y)

    def gradient_batch(self, sol, dprim, return_canonical=False):
        """dprim: dict name -> (B, *shape) or packed (B, n_prim) of the QP family; -> dict name -> (B, size) parameter gradients."""
        return self.qp.gradient_batch(sol.sol_y, dprim, return_canonical=return_canonical)
