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
from .blob import pack_blob, pack_tail_blob, pack_grad_blob
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
    refactor: Optional[RefactorTables] = None
    solve_source: str = ''
    theta_shared: Optional[np.ndarray] = None
    batch_cols: Optional[np.ndarray] = None
    stats: Dict[str, float] = field(default_factory=dict)


def setup_qp_family(fam: CanonFamily, batch_params: Optional[List[str]] = None,
                    theta: Optional[np.ndarray] = None, rho: float = 0.1, sigma: float = 1e-6,
                    scaling: int = 10, max_group_rows: int = 32, allow_trailing: bool = True) -> QPSetup:
    if fam.solver_type != 'quadratic':
        raise ValueError('ADMM-CUDA handles the QP canonical form only')
    if batch_params is None:
        batch_params = [p.name for p in fam.params if not (fam.changes('P', [p.name]) or fam.changes('A', [p.name]))]
    for name in batch_params:
        fam.param(name)      # AttributeError for unknown names, like the reference's cpg_solve
        if fam.changes('P', [name]) or fam.changes('A', [name]):
            raise ValueError(f'parameter {name} enters a canonical matrix; per-instance matrix updates '
                             'are not generated yet (shared parameters may: they are folded at setup)')
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
    F = _kkt.factorize(K, n)
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
    solve_source = pack_blob.last_solve_source
    # the tail kernel needs everything but the (large) tile schedule: a compact copy leaves room for more warps
    blob_compact = pack_blob(n=n, m=m, perm=F.perm, schedule=S, Ps_upper=sc['P'], As=sc['A'], D=sc['D'], E=sc['E'],
                             c=sc['c'], sigma=sigma, rho=rho, ctype=ctype, q_base=q_base, l_base=l_base, u_base=u_base,
                             Mq_b=Mq_b, Ml_b=Ml_b, Mu_b=Mu_b, npb=npb, prim_idx=prim_idx, dual_idx=dual_idx,
                             d_const=d_const, is_max=fam.is_maximization, with_tiles=False)
    RT = build_refactor_tables(F, K, n)
    tail_blob = pack_tail_blob(RT)
    # backward pass (gradient=True): regularised KKT of the UNSCALED problem on the same symbolic pattern
    grad_blob, grad_S0 = pack_grad_blob(n=n, m=m, perm=F.perm, P_upper=sp.csc_matrix(P), A=sp.csc_matrix(A), slot_of=RT.slot_of,
                                        n_slots=RT.n_slots, Mq_b=Mq_b, Ml_b=Ml_b, Mu_b=Mu_b, npb=npb, prim_idx=prim_idx)
    import struct as _struct
    from .blob import HEADER_FIELDS as _HF
    _hdr = dict(zip([n_ for _, n_ in _HF], _struct.unpack('<' + ''.join('i' if t_ == 'int' else 'd' for t_, _ in _HF),
                                                            blob[:_struct.calcsize('<' + ''.join('i' if t_ == 'int' else 'd' for t_, _ in _HF))])))
    st = dict(nb_slots=int(_hdr['nb_slots']), nnz_L=F.nnz, tail_blob_bytes=len(tail_blob), refactor_ops=len(RT.ops), n_levels=int(F.level.max()) + 1, n_tiles=len(S.tiles), schedule_cost=S.model_cost,
              schedule_entries=S.n_entries, blob_bytes=len(blob))
    return QPSetup(family=fam, batch_params=list(batch_params), n=n, m=m, npb=npb, rho=rho, sigma=sigma,
                   scaling=scaling, D=sc['D'], E=sc['E'], c=sc['c'], ctype=ctype, P_scaled=sc['P'],
                   A_scaled=sc['A'], factor=F, schedule=S, prim_idx=prim_idx, dual_idx=dual_idx, blob=blob, tail_blob=tail_blob, blob_compact=blob_compact, grad_blob=grad_blob, grad_S0=grad_S0, refactor=RT, solve_source=solve_source,
                   theta_shared=theta0, batch_cols=bcols, stats=st)
