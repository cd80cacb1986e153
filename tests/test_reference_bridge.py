"""f3 groundwork: `CanonFamily.from_reference_canon` against the reference's OWN dataclasses.

cvxpy is absent here, so the reference's `Canonicalizer` cannot run -- but its hand-off types (`cvxpygen/mappings.py`:
Canon, ParameterCanon, ParameterInfo, PrimalVariableInfo, DualVariableInfo) need only numpy / scipy and are loaded from
the reference tree by file path.  The test fills them the way `canonicalizer.py:124-332` does (user parameters in
user-sparsity column order with `flat_usp` ending in 1.0, `p_id_to_mapping` CSR per canonical id, `p` holding the canonical
matrices, duals as (vector name, indices)), runs the bridge, and checks that the family that comes out generates the very
same constants blob as the hand-built one.  Skipped where /root/reference does not exist (the GPU box)."""
import importlib.util
import os
from types import SimpleNamespace

import numpy as np
import pytest
import scipy.sparse as sp

from cvxpygen_b200 import families
from cvxpygen_b200.ir import CanonFamily
from cvxpygen_b200.offline.qp_setup import setup_qp_family

REF = os.environ.get('CPG_REFERENCE', '/root/reference')
MAPPINGS = os.path.join(REF, 'cvxpygen', 'mappings.py')
pytestmark = pytest.mark.skipif(not os.path.exists(MAPPINGS), reason='reference tree not present')


def _mappings():
    spec = importlib.util.spec_from_file_location('cvxpygen_ref_mappings', MAPPINGS)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _to_reference_canon(fam, M):
    """A `Canon` (reference dataclasses) carrying the family, filled like cvxpygen/canonicalizer.py does."""
    pinfo = M.ParameterInfo(
        col_to_name_usp={p.col: p.name for p in fam.params}, flat_usp=fam.theta_default(),
        id_to_col={i: p.col for i, p in enumerate(fam.params)}, ids=list(range(len(fam.params))),
        name_to_shape={p.name: p.shape for p in fam.params}, name_to_size_usp={p.name: p.size for p in fam.params},
        names=[p.name for p in fam.params], num=len(fam.params), writable={}, lower=None, upper=None)
    pcanon = M.ParameterCanon()
    pcanon.is_maximization = fam.is_maximization
    for pid, mp in fam.maps.items():
        pcanon.p_id_to_mapping[pid] = sp.csr_matrix(mp)
        pcanon.p_id_to_changes[pid] = bool(mp[:, :-1].nnz)
        pcanon.p[pid] = fam.canon_matrix(pid) if pid in fam.patterns else fam.canon_data(pid)
    off = 0
    n2o, n2i, n2s, n2sh, n2init, n2sym = {}, {}, {}, {}, {}, {}
    for v in fam.variables:
        n2o[v.name] = int(v.indices[0]); n2i[v.name] = v.indices; n2s[v.name] = len(v.indices); n2sh[v.name] = v.shape
        n2init[v.name] = np.zeros(v.shape); n2sym[v.name] = False
    pv = M.PrimalVariableInfo(n2o, n2i, n2s, n2sh, n2init, n2sym, [False] * len(fam.variables))
    dv = M.DualVariableInfo({d.name: int(d.indices[0]) for d in fam.duals}, {d.name: (d.vec, d.indices) for d in fam.duals},
                            {d.name: len(d.indices) for d in fam.duals}, {d.name: d.shape for d in fam.duals},
                            {d.name: np.zeros(len(d.indices)) for d in fam.duals}, {d.name: d.vec for d in fam.duals})
    return M.Canon(pv, dv, pinfo, pcanon)


@pytest.mark.parametrize('builder,batch', [(lambda: families.mpc(6, 3, 10), ['x_init']),
                                           (lambda: families.nonneg_ls(3, 2), ['b']),
                                           (lambda: families.mpc_reference(10), ['A', 'B', 'x_init'])])
def test_bridge_roundtrip_generates_identical_constants(builder, batch):
    M = _mappings()
    fam = builder()
    canon = _to_reference_canon(fam, M)
    iface = SimpleNamespace(solver_type='quadratic', n_var=fam.n_var, n_eq=fam.n_eq, n_ineq=fam.n_ineq)   # QPCanonMixin attributes
    fam2 = CanonFamily.from_reference_canon(fam.name, canon, iface)
    assert [p.name for p in fam2.params] == [p.name for p in fam.params]
    assert all(np.array_equal(a.default, b.default) and a.col == b.col and a.size == b.size for a, b in zip(fam.params, fam2.params))
    for pid in fam.maps:
        assert (fam.maps[pid] != fam2.maps[pid]).nnz == 0
    for pid in fam.patterns:
        assert all(np.array_equal(np.asarray(a), np.asarray(b)) for a, b in zip(fam.patterns[pid][:2], fam2.patterns[pid][:2]))
    assert [v.name for v in fam2.variables] == [v.name for v in fam.variables]
    assert all(np.array_equal(a.indices, b.indices) and a.vec == b.vec for a, b in zip(fam.duals, fam2.duals))
    s1, s2 = setup_qp_family(fam, batch), setup_qp_family(fam2, batch)
    assert s1.blob == s2.blob and s1.tail_blob == s2.tail_blob and s1.mat_blob == s2.mat_blob and s1.grad_blob == s2.grad_blob


def _class_level_names(path, cls):
    """names assigned / defined at class level of `cls` in a reference source file (parsed, not imported: cvxpy is absent)"""
    import ast
    tree = ast.parse(open(path).read())
    node = next(n for n in ast.walk(tree) if isinstance(n, ast.ClassDef) and n.name == cls)
    names = set()
    for st in node.body:
        if isinstance(st, ast.Assign):
            names |= {t.id for t in st.targets if isinstance(t, ast.Name)}
        elif isinstance(st, ast.AnnAssign) and isinstance(st.target, ast.Name):
            names.add(st.target.id)
        elif isinstance(st, ast.FunctionDef):
            names.add(st.name)
    return names


def test_plugin_classes_carry_the_reference_interface():
    """b1: ADMMCUDAInterface / IPMCUDAInterface expose every class-level attribute and method that the reference plugins they
    stand beside (OSQPInterface + QPCanonMixin, ECOSInterface) define -- except what only makes sense for a CPU solver whose
    workspace the writer addresses field by field (ws_ptrs) and the cvxpy-side affine-map plumbing the mixin provides."""
    from cvxpygen_b200.solvers import ADMMCUDAInterface, IPMCUDAInterface
    sol = os.path.join(REF, 'cvxpygen', 'solvers')
    osqp = _class_level_names(os.path.join(sol, 'osqp.py'), 'OSQPInterface') | \
        _class_level_names(os.path.join(sol, '_interface.py'), 'QPCanonMixin')
    ecos = _class_level_names(os.path.join(sol, 'ecos.py'), 'ECOSInterface')
    # the QP plugin carries EVERYTHING the writer reads (it is driven by the reference's own writer in tests/test_refwriter.py);
    # only the cvxpy-side affine-map plumbing, which the reference mixin provides when cvxpygen is importable, is left out
    cvxpy_side = {'get_affine_map', 'augment_vector_parameter', '__init__'}
    missing = {n for n in osqp - cvxpy_side if not hasattr(ADMMCUDAInterface, n)}
    assert not missing, sorted(missing)
    for n in ('ws_ptrs', 'declare_workspace', 'define_workspace', 'cmake_context_extra', 'setup_py_context', 'parameter_update_structure'):
        assert hasattr(ADMMCUDAInterface, n), n
    assert set(ADMMCUDAInterface.parameter_update_structure) == {'PA', 'P', 'A', 'qlu', 'ql', 'qu', 'lu', 'q', 'l', 'u'}   # osqp.py:20-61
    cpu_only = cvxpy_side | {'ws_ptrs', 'cmake_context_extra', 'setup_py_context', 'declare_workspace', 'define_workspace'}
    missing = {n for n in ecos - cpu_only if not hasattr(IPMCUDAInterface, n)}
    assert not missing, ('IPMCUDAInterface', sorted(missing))
    assert ADMMCUDAInterface.canon_p_ids == ['P', 'q', 'd', 'A', 'l', 'u'] and IPMCUDAInterface.canon_p_ids == ['c', 'd', 'A', 'b', 'G', 'h']
    assert IPMCUDAInterface.status_is_int and IPMCUDAInterface.dual_var_names == ['y', 'z'] and IPMCUDAInterface.dual_var_split
    i = IPMCUDAInterface(family=families.adp_socp())
    assert i.canon_constants['q'] == [8, 5, 4, 4] and i.stgs_translation == {'max_iters': 'maxit'}
    with pytest.raises(ValueError):
        IPMCUDAInterface.check_unsupported_cones(SimpleNamespace(exp=1))
