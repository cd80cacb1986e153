"""Generation-time setup of one QP family for the ADMM-CUDA backend (the offline half of a3/a10).

Mirrors what `osqp_setup` does once per problem (osqp_sources/src/osqp.c:76-283):
scale data, classify constraints / build rho_vec, form + order + factor the KKT
matrix -- and then re-expresses the factor as a warp schedule and packs the constants blob.
"""
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np
import scipy.sparse as sp

from ..ir import CanonFamily
from . import kkt as _kkt
from .blob import pack_blob, pack_tail_blob, pack_grad_blob, pack_matpar_blob, last_solve_source as _last_solve_source
from .refactor import build_refactor_tables, RefactorTables
from .equilibrate import ruiz_equilibrate
from .schedule import SolveSchedule, build_schedule

OSQP_INFTY = 1e30


@dataclass
class QPSetup:
    family: CanonFamily
    batch_params: List[str]
    n: int
    m: int
    npb: int
    rho: float
    sigma: float
    scaling: int
    D: np.ndarray
    E: np.ndarray
    c: float
    ctype: np.ndarray
    P_scaled: sp.csc_matrix
    A_scaled: sp.csc_matrix
    factor: _kkt.LDLFactor
    schedule: SolveSchedule
    prim_idx: np.ndarray
    dual_idx: np.ndarray
    blob: bytes = b''
    tail_blob: bytes = b''
    blob_compact: bytes = b''
    grad_blob: bytes = b''
    grad_S0: Optional[np.ndarray] = None
    mat_blob: bytes = b''                  # f2: per-instance matrix parameters (empty = matrices shared by the batch)
    mat_params: List[str] = field(default_factory=list)
    nnzP: int = 0
    nnzA: int = 0
    refactor: Optional[RefactorTables] = None
    solve_source: str = ''
    theta_shared: Optional[np.ndarray] = None
    batch_cols: Optional[np.ndarray] = None
    stats: Dict[str, float] = field(default_factory=dict)
    max_group_rows: int = 32
    allow_trailing: bool = True
    dmma_blob: bytes = b''                 # tables of the FP64 tensor-core solve (offline/dmma.py); empty = not generated
    dmma_rounds: int = 0


def setup_qp_family(fam: CanonFamily, batch_params: Optional[List[str]] = None,
                    theta: Optional[np.ndarray] = None, rho: float = 0.1, sigma: float = 1e-6,
                    scaling: int = 10, max_group_rows: int = 32, allow_trailing: bool = True, dmma: bool = True) -> QPSetup:
    if fam.solver_type != 'quadratic':
        raise ValueError('ADMM-CUDA handles the QP canonical form only')
    if batch_params is None:
        batch_params = [p.name for p in fam.params if not (fam.changes('P', [p.name]) or fam.changes('A', [p.name]))]
    mat_params = []
    for name in batch_params:
        fam.param(name)      # AttributeError for unknown names, like the reference's cpg_solve
        if fam.changes('P', [name]) or fam.changes('A', [name]):
            mat_params.append(name)      # f2: this family is solved by the per-instance matrix kernel (matpar_kernel.cuh)
    theta = fam.theta_default() if theta is None else np.asarray(theta, dtype=float)
    n, m = fam.n_var, fam.n_eq + fam.n_ineq
    bcols = fam.param_columns(batch_params) if batch_params else np.zeros(0, dtype=int)
    npb = len(bcols)
    P = fam.canon_matrix('P', theta) if 'P' in fam.maps else sp.csc_matrix((n, n))
    A = fam.canon_matrix('A', theta)
    q = fam.canon_data('q', theta)
    l = np.clip(fam.canon_data('l', theta), -OSQP_INFTY, OSQP_INFTY)
    u = np.clip(fam.canon_data('u', theta), -OSQP_INFTY, OSQP_INFTY)
    if scaling:
        sc = ruiz_equilibrate(P, A, q, scaling)
    else:
        sc = dict(P=sp.csc_matrix(P, dtype=float), A=sp.csc_matrix(A, dtype=float), q=q.copy(),
                  D=np.ones(n), E=np.ones(m), c=1.0)
    ctype = _kkt.constraint_types(sc['E'] * l, sc['E'] * u)
    rho = min(max(rho, _kkt.RHO_MIN), _kkt.RHO_MAX)
    rho_vec = _kkt.rho_vector(ctype, rho)
    K = _kkt.assemble_kkt(sc['P'], sc['A'], sigma, rho_vec)
    Kpat = None
    if mat_params:
        # batched matrix entries may take any value: order / analyse the STRUCTURAL pattern, not today's nonzeros
        ones = lambda M: sp.csc_matrix((np.ones(len(M.indices)), M.indices, M.indptr), shape=M.shape)
        Kpat = _kkt.assemble_kkt(ones(sp.csc_matrix(P)), ones(sp.csc_matrix(A)), 1.0, np.ones(m))
    F = _kkt.factorize(K, n, pattern=Kpat)
    S = build_schedule(F, max_group_rows=max_group_rows, allow_trailing=allow_trailing)
    # affine maps split into [batched columns | everything else folded into a base vector]
    theta0 = theta.copy(); theta0[bcols] = 0.0

    def split(pid, rows):
        M = fam.maps.get(pid)
        if M is None:
            return np.zeros(rows), sp.csr_matrix((rows, max(npb, 1)))
        base = np.asarray(M @ theta0).ravel()
        Mb = sp.csr_matrix(M[:, bcols]) if npb else sp.csr_matrix((rows, 1))
        return base, Mb
    q_base, Mq_b = split('q', n)
    l_base, Ml_b = split('l', m)
    u_base, Mu_b = split('u', m)
    l_base = np.clip(l_base, -OSQP_INFTY, OSQP_INFTY)
    u_base = np.clip(u_base, -OSQP_INFTY, OSQP_INFTY)
    d_const = float(np.asarray(fam.maps['d'] @ theta).ravel()[0]) if 'd' in fam.maps else 0.0
    if 'd' in fam.maps and npb and fam.maps['d'][:, bcols].nnz:
        raise ValueError('objective offset d depending on a batched parameter is not generated yet')
    prim_idx = np.concatenate([v.indices for v in fam.variables]) if fam.variables else np.zeros(0, int)
    dual_idx = np.concatenate([d.indices for d in fam.duals]) if fam.duals else np.zeros(0, int)
    blob = pack_blob(n=n, m=m, perm=F.perm, schedule=S, Ps_upper=sc['P'], As=sc['A'], D=sc['D'], E=sc['E'],
                     c=sc['c'], sigma=sigma, rho=rho, ctype=ctype, q_base=q_base, l_base=l_base, u_base=u_base,
                     Mq_b=Mq_b, Ml_b=Ml_b, Mu_b=Mu_b, npb=npb, prim_idx=prim_idx, dual_idx=dual_idx,
                     d_const=d_const, is_max=fam.is_maximization)
    solve_source = _last_solve_source()
    # the tail kernel needs everything but the (large) tile schedule: a compact copy leaves room for more warps
    blob_compact = pack_blob(n=n, m=m, perm=F.perm, schedule=S, Ps_upper=sc['P'], As=sc['A'], D=sc['D'], E=sc['E'],
                             c=sc['c'], sigma=sigma, rho=rho, ctype=ctype, q_base=q_base, l_base=l_base, u_base=u_base,
                             Mq_b=Mq_b, Ml_b=Ml_b, Mu_b=Mu_b, npb=npb, prim_idx=prim_idx, dual_idx=dual_idx,
                             d_const=d_const, is_max=fam.is_maximization, with_tiles=False)
    dmma_blob, dmma_rounds = b'', 0
    if dmma and not mat_params and len(blob) <= 227 * 1024:      # (families beyond one SM's shared memory take the per-instance-factor route)
        # shared KKT factor: the solve of eight instances is a dense contraction -> FP64 tensor cores
        from .dmma import build_dmma_schedule, pack_dmma_blob
        try:
            DS = build_dmma_schedule(F, max_group_rows=max_group_rows)
            dmma_blob = pack_dmma_blob(DS)
            dmma_rounds = max(len(t.round_len) for t in DS.tiles)
        except AssertionError:      # 16-bit operand offsets: no tensor-core schedule for a family of this size
            dmma_blob, dmma_rounds = b'', 0
    RT = build_refactor_tables(F, K, n)
    tail_blob = pack_tail_blob(RT)
    # backward pass (gradient=True): regularised KKT of the UNSCALED problem on the same symbolic pattern
    gkw = {}
    if mat_params:      # rows a16 + f2: transposed maps of the P / A entries -> dtheta of the matrix parameters
        gkw = dict(MP_b=sp.csr_matrix(fam.maps['P'][:, bcols]), MA_b=sp.csr_matrix(fam.maps['A'][:, bcols]))
    grad_blob, grad_S0 = pack_grad_blob(n=n, m=m, perm=F.perm, P_upper=sp.csc_matrix(P), A=sp.csc_matrix(A), slot_of=RT.slot_of,
                                        n_slots=RT.n_slots, Mq_b=Mq_b, Ml_b=Ml_b, Mu_b=Mu_b, npb=npb, prim_idx=prim_idx, **gkw)
    mat_blob = b''
    if mat_params:
        mat_blob = _matpar_blob(fam, sc, P, A, q, theta0, bcols, npb, F, RT, sigma, scaling, n, m)
    import struct as _struct
    from .blob import HEADER_FIELDS as _HF
    _hdr = dict(zip([n_ for _, n_ in _HF], _struct.unpack('<' + ''.join('i' if t_ == 'int' else 'd' for t_, _ in _HF),
                                                            blob[:_struct.calcsize('<' + ''.join('i' if t_ == 'int' else 'd' for t_, _ in _HF))])))
    st = dict(nb_slots=int(_hdr['nb_slots']), nnz_L=F.nnz, tail_blob_bytes=len(tail_blob), refactor_ops=len(RT.ops), n_levels=int(F.level.max()) + 1, n_tiles=len(S.tiles), schedule_cost=S.model_cost,
              schedule_entries=S.n_entries, blob_bytes=len(blob), dmma_blob_bytes=len(dmma_blob))
    return QPSetup(family=fam, batch_params=list(batch_params), n=n, m=m, npb=npb, rho=rho, sigma=sigma,
                   scaling=scaling, D=sc['D'], E=sc['E'], c=sc['c'], ctype=ctype, P_scaled=sc['P'],
                   A_scaled=sc['A'], factor=F, schedule=S, prim_idx=prim_idx, dual_idx=dual_idx, blob=blob, tail_blob=tail_blob, blob_compact=blob_compact, grad_blob=grad_blob, grad_S0=grad_S0, mat_blob=mat_blob, mat_params=mat_params,
                   nnzP=int(sp.csc_matrix(P).indptr[-1]), nnzA=int(sp.csc_matrix(A).indptr[-1]), refactor=RT, solve_source=solve_source,
                   theta_shared=theta0, batch_cols=bcols, stats=st,
                   max_group_rows=max_group_rows, allow_trailing=allow_trailing, dmma_blob=dmma_blob, dmma_rounds=dmma_rounds)


def unscale_roundtrip(sc, scaling):
    """What `unscale_data` (scaling.c:160-175) leaves in the workspace when osqp_update_P_A starts from the pristine
    post-setup state: P, q, A un-scaled from their SCALED values in the reference's operation order (not bit-identical
    to the original data).  Returns (P.data, A.data, q)."""
    P, A = sp.csc_matrix(sc['P']), sp.csc_matrix(sc['A'])
    if not scaling:
        return P.data.copy(), A.data.copy(), np.array(sc['q'], dtype=float)
    D, E, c = sc['D'], sc['E'], sc['c']
    Dinv, Einv, cinv = 1.0 / D, 1.0 / E, 1.0 / c
    n = P.shape[0]
    pc = np.repeat(np.arange(n), np.diff(P.indptr)); ac = np.repeat(np.arange(n), np.diff(A.indptr))
    Pd = ((P.data * cinv) * Dinv[P.indices]) * Dinv[pc]
    qd = Dinv * (np.asarray(sc['q'], dtype=float) * cinv)
    Ad = (A.data * Einv[A.indices]) * Dinv[ac]
    return Pd, Ad, qd


def _matpar_blob(fam, sc, P, A, q, theta0, bcols, npb, F, RT, sigma, scaling, n, m):
    P, A = sp.csc_matrix(P), sp.csc_matrix(A)
    Pd_un, Ad_un, q_un = unscale_roundtrip(sc, scaling)
    MP, MA = fam.maps['P'], fam.maps['A']
    P_batched = MP[:, bcols].nnz > 0
    A_batched = MA[:, bcols].nnz > 0
    # osqp_update_data_mat is called with the WHOLE canonical P->x / A->x of the matrices that are outdated and with NULL
    # for the other one (cvxpygen/solvers/osqp.py:20-33), which then keeps its unscale/re-scale round-trip values
    P_base = np.asarray(MP @ theta0).ravel() if P_batched else Pd_un
    A_base = np.asarray(MA @ theta0).ravel() if A_batched else Ad_un
    MP_b = sp.csr_matrix(MP[:, bcols]) if P_batched else sp.csr_matrix((MP.shape[0], max(npb, 1)))
    MA_b = sp.csr_matrix(MA[:, bcols]) if A_batched else sp.csr_matrix((MA.shape[0], max(npb, 1)))
    return pack_matpar_blob(n=n, m=m, perm=F.perm, P_pattern=fam.patterns['P'], A_pattern=fam.patterns['A'],
                            slot_of=RT.slot_of, n_slots=RT.n_slots, sigma=sigma, MP_b=MP_b, MA_b=MA_b, P_base=P_base,
                            A_base=A_base, q_un=q_un, npb=npb, scaling_iters=scaling)
