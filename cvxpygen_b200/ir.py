"""Solver-independent intermediate representation of one DPP problem family.

This is the cvxpy-independent hand-off between canonicalisation and the CUDA
code generator.  It carries the same information the reference keeps in its
``Canon`` bundle (reference: cvxpygen/mappings.py:34-145 -- ParameterCanon,
ParameterInfo, PrimalVariableInfo, DualVariableInfo) but organised per object
instead of per attribute, so that a family can be built either

  * by hand, without cvxpy (``cvxpygen_b200.families``), or
  * from the reference's own ``Canonicalizer`` output when cvxpy is installed
    (``CanonFamily.from_reference_canon``).

Conventions (reference: SURVEY Appendix B; cvxpygen/canonicalizer.py:226-332):

  theta = [user parameters flattened in user sparsity, Fortran order ; 1.0]
  canonical object ``id``  =  maps[id] @ theta            (CSR, one row per stored entry)

QP form   (reference: cvxpygen/solvers/_interface.py:18-79):
      min 1/2 x'Px + q'x + d   s.t.  l <= Ax <= u,  rows = [equalities ; inequalities]
      P is the upper-triangular CSC, A is CSC; the maps produce their ``.data``.
Conic form (reference: cvxpygen/solvers/_interface.py:132-173):
      min c'x + d  s.t.  Ax = b,  h - Gx in K.
"""
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np
import scipy.sparse as sp


@dataclass
class UserParam:
    """One user-level parameter (reference: ParameterInfo, cvxpygen/mappings.py:51-68)."""
    name: str
    shape: Tuple[int, ...]
    size: int            # number of stored entries (user sparsity: diag / sparse params store fewer)
    col: int             # first column of this parameter in theta
    default: np.ndarray  # flat default value, length ``size``


@dataclass
class UserVar:
    """One user-level primal variable: a gather from the canonical solution
    (reference: PrimalVariableInfo.name_to_indices, cvxpygen/canonicalizer.py:124-158)."""
    name: str
    shape: Tuple[int, ...]
    indices: np.ndarray  # positions in canonical x, Fortran order of the variable


@dataclass
class UserDual:
    """One user-level constraint dual (reference: DualVariableInfo.name_to_indices,
    cvxpygen/canonicalizer.py:160-224).  ``vec`` is 'y' for the QP form, 'y'/'z' for conic."""
    name: str
    vec: str
    shape: Optional[Tuple[int, ...]]
    indices: np.ndarray


@dataclass
class CanonFamily:
    name: str
    solver_type: str                      # 'quadratic' | 'conic'
    n_var: int
    n_eq: int
    n_ineq: int
    params: List[UserParam]
    maps: Dict[str, sp.csr_matrix]        # canonical id -> CSR (entries x n_theta)
    patterns: Dict[str, Tuple[np.ndarray, np.ndarray, Tuple[int, int]]]  # 'P','A','G' -> (indices, indptr, shape) CSC
    variables: List[UserVar]
    duals: List[UserDual]
    is_maximization: bool = False
    cone_dims: Dict[str, object] = field(default_factory=dict)   # conic: {'l': int, 'q': [int,...]}

    # ---- parameter vector ---------------------------------------------------
    @property
    def n_theta(self) -> int:
        return sum(p.size for p in self.params) + 1

    def param(self, name: str) -> UserParam:
        for p in self.params:
            if p.name == name:
                return p
        raise AttributeError(f'{name} is not a parameter.')   # same error text as TPL/cpg_solver.py.jinja2:50-51

    def theta_default(self) -> np.ndarray:
        th = np.ones(self.n_theta)
        for p in self.params:
            th[p.col:p.col + p.size] = np.asarray(p.default, dtype=float).ravel()
        return th

    # ---- canonical data -----------------------------------------------------
    def canon_data(self, p_id: str, theta: Optional[np.ndarray] = None) -> np.ndarray:
        """Stored entries of canonical object ``p_id`` for one theta (a2: cpg_canonicalize_<id>,
        reference emitter cvxpygen/utils.py:279-294)."""
        th = self.theta_default() if theta is None else theta
        return np.asarray(self.maps[p_id] @ th).ravel()

    def canon_matrix(self, p_id: str, theta: Optional[np.ndarray] = None) -> sp.csc_matrix:
        idx, ptr, shape = self.patterns[p_id]
        return sp.csc_matrix((self.canon_data(p_id, theta), idx.copy(), ptr.copy()), shape=shape)

    def changes(self, p_id: str, names: Optional[List[str]] = None) -> bool:
        """Does canonical object p_id depend on (the given subset of) user parameters?
        (reference: p_id_to_changes, cvxpygen/canonicalizer.py:324)."""
        M = self.maps.get(p_id)
        if M is None:
            return False
        cols = self.param_columns(names)
        return M[:, cols].nnz > 0 if len(cols) else False

    def param_columns(self, names: Optional[List[str]] = None) -> np.ndarray:
        ps = self.params if names is None else [self.param(n) for n in names]
        if not ps:
            return np.zeros(0, dtype=int)
        return np.concatenate([np.arange(p.col, p.col + p.size) for p in ps])

    def outdated_by(self, name: str) -> List[str]:
        """canonical ids touched by one user parameter (adjacency, cvxpygen/canonicalizer.py:117-120)."""
        return [k for k in self.maps if self.changes(k, [name])]

    # ---- a family straight from canonical QP data (no modelling layer at all) -----------------------------------
    @classmethod
    def from_canonical_qp(cls, name, P, q, A, l, u, n_eq=None, matrix_params=False) -> 'CanonFamily':
        """min 1/2 x'Px + q'x  s.t.  l <= Ax <= u  with the canonical vectors themselves as user parameters ``q``, ``l``,
        ``u`` (and, with ``matrix_params``, the stored entries of ``P`` (upper triangle) and ``A`` in CSC order as ``P``,
        ``A``): the identity maps the reference would emit for a problem whose parameters ARE its canonical data.
        Structural zeros of P / A stay in the pattern.  Variables: ``x``; duals: ``y``."""
        P = sp.triu(sp.csc_matrix(P), format='csc'); P.sort_indices()
        A = sp.csc_matrix(A); A.sort_indices()
        n, m = P.shape[0], A.shape[0]
        q = np.asarray(q, dtype=float).ravel(); l = np.asarray(l, dtype=float).ravel(); u = np.asarray(u, dtype=float).ravel()
        specs = [('q', (n,), q), ('l', (m,), l), ('u', (m,), u)]
        mats = ('P', 'A') if matrix_params is True else tuple(matrix_params or ())     # True = both; or a subset, e.g. ('A',)
        if 'P' in mats:
            specs.append(('P', (P.nnz,), P.data))
        if 'A' in mats:
            specs.append(('A', (A.nnz,), A.data))
        params, col = [], 0
        for nm, shape, default in specs:
            params.append(UserParam(nm, tuple(shape), int(np.asarray(default).size), col, np.array(default, dtype=float)))
            col += params[-1].size
        n_theta = col + 1
        ident = lambda size, c0: sp.csr_matrix((np.ones(size), (np.arange(size), c0 + np.arange(size))), shape=(size, n_theta))
        const = lambda v: sp.csr_matrix((np.asarray(v, dtype=float), (np.arange(len(v)), np.full(len(v), n_theta - 1))),
                                        shape=(len(v), n_theta))
        pc = {p.name: p.col for p in params}
        maps = {'q': ident(n, pc['q']), 'l': ident(m, pc['l']), 'u': ident(m, pc['u']), 'd': sp.csr_matrix((1, n_theta)),
                'P': ident(P.nnz, pc['P']) if 'P' in mats else const(P.data),
                'A': ident(A.nnz, pc['A']) if 'A' in mats else const(A.data)}
        if n_eq is None:
            n_eq = int(np.sum(l == u))
        pat = lambda M: (M.indices.astype(np.int32), M.indptr.astype(np.int32), M.shape)
        return cls(name, 'quadratic', n, n_eq, m - n_eq, params, maps, {'P': pat(P), 'A': pat(A)},
                   [UserVar('x', (n,), np.arange(n))], [UserDual('y', 'y', (m,), np.arange(m))])

    @classmethod
    def from_canonical_conic(cls, name, c, A, b, G, h, l, q=(), matrix_params=()) -> 'CanonFamily':
        """min c'x  s.t.  Ax = b,  h - Gx in R+^l x SOC(q_1) x ...  (ECOS form) with the canonical vectors ``c``, ``b``, ``h`` as
        the user parameters and A, G constant: the conic counterpart of from_canonical_qp.  Variables: ``x``; duals ``y`` (if
        there are equalities) and ``z``."""
        A = sp.csc_matrix(A); A.sort_indices(); G = sp.csc_matrix(G); G.sort_indices()
        n, p, m = G.shape[1], A.shape[0], G.shape[0]
        specs = [('c', (n,), c)] + ([('b', (p,), b)] if p else []) + [('h', (m,), h)]
        mats = ('G', 'A') if matrix_params is True else tuple(matrix_params or ())    # stored entries (CSC order) as parameters ``G`` / ``A``
        if 'G' in mats:
            specs.append(('G', (G.nnz,), G.data))
        if 'A' in mats and p:
            specs.append(('A', (A.nnz,), A.data))
        params, col = [], 0
        for nm, shape, default in specs:
            d = np.asarray(default, dtype=float).ravel()
            params.append(UserParam(nm, tuple(shape), int(d.size), col, d.copy()))
            col += d.size
        n_theta = col + 1
        pc = {pp.name: pp.col for pp in params}
        ident = lambda size, c0: sp.csr_matrix((np.ones(size), (np.arange(size), c0 + np.arange(size))), shape=(size, n_theta))
        const = lambda v: sp.csr_matrix((np.asarray(v, dtype=float), (np.arange(len(v)), np.full(len(v), n_theta - 1))),
                                        shape=(len(v), n_theta))
        maps = {'c': ident(n, pc['c']), 'b': ident(p, pc['b']) if p else sp.csr_matrix((0, n_theta)), 'h': ident(m, pc['h']),
                'd': sp.csr_matrix((1, n_theta)), 'A': ident(A.nnz, pc['A']) if 'A' in pc else const(A.data),
                'G': ident(G.nnz, pc['G']) if 'G' in pc else const(G.data)}
        pat = lambda M: (M.indices.astype(np.int32), M.indptr.astype(np.int32), M.shape)
        duals = ([UserDual('y', 'y', (p,), np.arange(p))] if p else []) + [UserDual('z', 'z', (m,), np.arange(m))]
        return cls(name, 'conic', n, p, m, params, maps, {'A': pat(A), 'G': pat(G)}, [UserVar('x', (n,), np.arange(n))], duals,
                   cone_dims={'l': int(l), 'q': [int(d) for d in q]})

    # ---- bridge from the reference's own canonicaliser (exercised on the reference's dataclasses in tests/test_reference_bridge.py;
    #      the cvxpy-driven end of it cannot run in this image) ----
    @classmethod
    def from_reference_canon(cls, name, canon, solver_interface) -> 'CanonFamily':
        """Build the IR from (Canon, SolverInterface) as returned by the reference's
        ``Canonicalizer.canonicalize`` (cvxpygen/canonicalizer.py:47-52).  Only attribute
        reads -- nothing of cvxpygen is imported here."""
        pi, pc = canon.parameter_info, canon.parameter_canon
        pv, dv = canon.prim_variable_info, canon.dual_variable_info
        params = []
        for col in sorted(pi.col_to_name_usp):
            nm = pi.col_to_name_usp[col]
            sz = pi.name_to_size_usp[nm]
            params.append(UserParam(nm, tuple(pi.name_to_shape[nm]), sz, col,
                                    np.asarray(pi.flat_usp[col:col + sz], dtype=float)))
        maps = {k: sp.csr_matrix(v) for k, v in pc.p_id_to_mapping.items() if v is not None}
        patterns = {}
        for k, M in pc.p.items():
            if sp.issparse(M):
                M = sp.csc_matrix(M)
                patterns[k] = (M.indices.copy(), M.indptr.copy(), M.shape)
        variables = [UserVar(n, tuple(pv.name_to_shape[n]), np.asarray(pv.name_to_indices[n]))
                     for n in pv.name_to_indices]
        duals = [UserDual(n, v, dv.name_to_shape[n], np.asarray(ix))
                 for n, (v, ix) in dv.name_to_indices.items()]
        fam = cls(name, solver_interface.solver_type, solver_interface.n_var, solver_interface.n_eq,
                  solver_interface.n_ineq, params, maps, patterns, variables, duals,
                  is_maximization=pc.is_maximization)
        return fam
