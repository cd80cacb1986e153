"""Boundary rows b1-b3 / f3: the REFERENCE's own writer drives the ADMM-CUDA plugin.

CPU part (runs where /root/reference exists): `refwriter.write_reference_layout` calls the reference's emitters
(cvxpygen/utils.py: write_workspace_def/prot, write_solve_def/prot, write_module_def, cpg_module.hpp.jinja2), loaded from the
reference tree by path, with `ADMMCUDAInterface`; the emitted cpg_workspace.c / cpg_solve.c compile with gcc, the emitted
cpg_module.cpp with g++ + pybind11, everything links against the plugin's libcpg_b200.so, the module imports and exposes the
reference's API (solve, set_solver_*, cpg_params, cpg_updated, ...) plus the new `solve_batch`.
GPU part: `cpg_module.solve(upd, par)` through that PYBIND module equals the oracle (identical iterations, 1e-5); the ctypes
view of the emitted C ABI (`cpg_update_<p>`, `cpg_solve`, `CPG_Result`) does too; `solve_batch` equals a loop of `solve`."""
import ctypes as C
import os
import sysconfig

import numpy as np
import pytest

from cvxpygen_b200 import families, refwriter, standard
from helpers import canon_batches, oracle_solve

LAYOUTS = refwriter.standard_layouts()


def _dir(name):
    d = os.path.join(standard.GENERATED_DIR, name)
    ext = os.path.join(d, 'cpg_module' + sysconfig.get_config_var('EXT_SUFFIX'))
    if not os.path.exists(ext):
        if not refwriter.reference_available():
            pytest.skip('reference-layout directory not built and the reference tree is absent')
        refwriter.build_standard_layouts()
    return d


@pytest.mark.skipif(not refwriter.reference_available(), reason='reference tree not present')
def test_reference_writer_emits_compilable_code_for_the_plugin(tmp_path):
    """generation only (no nvcc): the decision tree and the workspace the reference writer emits for our attribute values"""
    fam = families.nonneg_ls(3, 2, name='nonneg_LS_3_2_A')
    d = str(tmp_path / 'code')
    canon, iface, cfg = refwriter.write_reference_layout(fam, d, prefix='7up')
    assert cfg.prefix == '_7up_'                                        # generator.py:175-182
    solve_c = open(os.path.join(d, 'c', 'src', 'cpg_solve.c')).read()
    ws_h = open(os.path.join(d, 'c', 'include', 'cpg_workspace.h')).read()
    ws_c = open(os.path.join(d, 'c', 'src', 'cpg_workspace.c')).read()
    # A and l/u are outdated by the user parameters A and b: OSQP's table restated with the shim's functions
    assert '_7up_cpg_b200_shim_update_mat(0, _7up_Canon_Params.A->x)' in solve_c
    assert '_7up_cpg_b200_shim_update_vec(0, _7up_Canon_Params.l, _7up_Canon_Params.u)' in solve_c
    assert '_7up_cpg_b200_shim_solve();' in solve_c and 'osqp' not in solve_c.lower()
    assert 'void _7up_cpg_update_A(cpg_int idx, cpg_float val)' in solve_c and 'void _7up_cpg_canonicalize_A()' in solve_c
    assert '(&_7up_cpg_b200_shim_settings)->eps_abs = eps_abs_new;' in solve_c
    assert '#include "cpg_b200_shim.h"' in ws_h and 'extern CpgB200ShimInfo _7up_cpg_b200_shim_info;' in ws_h
    assert 'cpg_float _7up_sol_x[5];' in ws_c and '&_7up_sol_x + 0' in ws_c            # CPG_Prim points into the shim's solution
    # the emitted C compiles as C99 against the shim header
    import subprocess
    inc, sol = os.path.join(d, 'c', 'include'), os.path.join(d, 'c', 'solver_code')
    for c in ('cpg_workspace.c', 'cpg_solve.c'):
        r = subprocess.run(['gcc', '-std=c99', '-fsyntax-only', '-I', inc, '-I', sol, os.path.join(d, 'c', 'src', c)], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    r = subprocess.run(['gcc', '-std=c99', '-fsyntax-only', '-I', sol, os.path.join(sol, 'cpg_b200_shim.c')], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    mod_cpp = open(os.path.join(d, 'cpp', 'src', 'cpg_module.cpp')).read()
    assert 'PYBIND11_MODULE(cpg_module, m)' in mod_cpp and '_7up_cpg_b200_register_batch(m);' in mod_cpp
    # build-system hooks the reference's templates read (cmake / setup.py contexts)
    assert 'cmake_target_link_libs' in iface.cmake_context_extra() and 'extra_objects' in iface.setup_py_context()


def test_pybind_module_loads_and_exposes_the_reference_api():
    d = _dir('refwriter_mpc_6_3_10')
    m = refwriter.load_module(d)
    for name in ('solve', 'solve_batch', 'cpg_params', 'cpg_updated', 'cpg_prim', 'cpg_dual', 'cpg_info', 'cpg_result',
                 'set_solver_default_settings', 'set_solver_max_iter', 'set_solver_eps_abs', 'set_solver_warm_starting'):
        assert hasattr(m, name), name
    par = m.cpg_params(); upd = m.cpg_updated()
    par.x_init = [0.0] * 6; upd.x_init = True
    with pytest.raises(AttributeError):
        m.solve_batch({'no_such_parameter': np.zeros((2, 6))})
    lib = C.CDLL(os.path.join(d, 'libcpg_b200.so'))          # the plugin's library next to it exports the batched C ABI
    assert hasattr(lib, 'cpg_solve_batch_host') and hasattr(lib, 'cpg_b200_kernel_times')


@pytest.mark.gpu
def test_pybind_solve_equals_oracle():
    """b3: cpg_module.solve(upd, par) through the reference-emitted pybind module (the 'lu' branch of the update tree)."""
    d = _dir('refwriter_mpc_6_3_10')
    m = refwriter.load_module(d)
    fam = families.mpc(6, 3, 10)
    rng = np.random.default_rng(4)
    m.set_solver_default_settings()
    m.set_solver_warm_starting(0)                 # every solve from the cold start, like the oracle
    for _ in range(4):
        xi = rng.uniform(-1, 1, 6)
        par = m.cpg_params(); upd = m.cpg_updated()
        par.x_init = list(xi); upd.x_init = True
        res = m.solve(upd, par)
        q, l, u = canon_batches(fam, {'x_init': xi[None, :]}, 1)
        ora = oracle_solve(fam, q, l, u)
        assert res.cpg_info.status == 'solved' and res.cpg_info.iter == int(ora['iter'][0])
        for v in fam.variables:
            got = np.asarray(getattr(res.cpg_prim, v.name))
            assert np.allclose(got, ora['x'][0, v.indices], rtol=1e-5, atol=1e-9), v.name
        for dv in fam.duals:
            assert np.allclose(np.asarray(getattr(res.cpg_dual, dv.name)), ora['y'][0, dv.indices], rtol=1e-5, atol=1e-9), dv.name
        assert abs(res.cpg_info.obj_val - ora['obj'][0]) < 1e-8
    # settings reach the kernel through cpg_set_solver_<name> -> shim settings
    m.set_solver_max_iter(25)
    res = m.solve(upd, par)
    assert res.cpg_info.iter == 25
    m.set_solver_default_settings()
    # warm start (OSQP's default for successive solves): the second solve of the same instance stops at the first check
    m.solve(upd, par)
    res2 = m.solve(upd, par)
    assert res2.cpg_info.iter == 25 and res2.cpg_info.status == 'solved'


@pytest.mark.gpu
def test_pybind_solve_batch_equals_oracle_and_single_solves():
    d = _dir('refwriter_mpc_6_3_10')
    m = refwriter.load_module(d)
    fam = families.mpc(6, 3, 10)
    B = 300
    xi = np.random.default_rng(8).uniform(-1, 1, (B, 6))
    out = m.solve_batch({'x_init': xi})
    q, l, u = canon_batches(fam, {'x_init': xi}, B)
    ora = oracle_solve(fam, q, l, u)
    assert np.array_equal(out['cpg_info']['iter'], ora['iter']) and np.array_equal(out['cpg_info']['status'], ora['status'])
    for v in fam.variables:
        assert np.allclose(out['cpg_prim'][v.name], ora['x'][:, v.indices], rtol=1e-5, atol=1e-9)
    for dv in fam.duals:
        assert np.allclose(out['cpg_dual'][dv.name], ora['y'][:, dv.indices], rtol=1e-5, atol=1e-9)
    assert np.allclose(out['cpg_info']['obj_val'], ora['obj'], rtol=1e-6, atol=1e-9)


@pytest.mark.gpu
def test_emitted_c_abi_via_ctypes_matrix_family_with_prefix():
    """b2 through ctypes: the symbols the reference writer emitted -- <p>cpg_update_<param>, <p>cpg_solve, <p>CPG_Result -- on
    the family whose sparse matrix A is a user parameter (update_mat branch -> per-instance re-equilibration + refactorisation on
    the GPU), generated with a prefix."""
    d = _dir('refwriter_nonneg_LS_3_2_A')
    ext = os.path.join(d, 'cpg_module' + sysconfig.get_config_var('EXT_SUFFIX'))
    lib = C.CDLL(ext)                          # the extension carries the emitted C objects
    fam = families.nonneg_ls(3, 2, name='nonneg_LS_3_2_A')

    class Prim(C.Structure):
        _fields_ = [(v.name, C.POINTER(C.c_double) if len(v.indices) > 1 else C.c_double) for v in fam.variables]

    class Dual(C.Structure):
        _fields_ = [(dv.name, C.POINTER(C.c_double) if len(dv.indices) > 1 else C.c_double) for dv in fam.duals]

    class Info(C.Structure):
        _fields_ = [('obj_val', C.c_double), ('iter', C.c_int), ('status', C.c_char_p), ('pri_res', C.c_double), ('dua_res', C.c_double)]

    class Result(C.Structure):
        _fields_ = [('prim', C.POINTER(Prim)), ('dual', C.POINTER(Dual)), ('info', C.POINTER(Info))]
    p = 'nnls_'
    fn = lambda name: getattr(lib, p + name)          # getattr caches the function object (lib[name] does not): argtypes stick
    fn('cpg_update_A').argtypes = [C.c_int, C.c_double]
    fn('cpg_update_b').argtypes = [C.c_int, C.c_double]
    fn('cpg_set_solver_warm_starting').argtypes = [C.c_int]
    fn('cpg_set_solver_default_settings')()
    fn('cpg_set_solver_warm_starting')(0)
    rng = np.random.default_rng(12)
    result = Result.in_dll(lib, p + 'CPG_Result')
    from helpers import canon_matrix_batches, matrix_oracle_solve
    for _ in range(3):
        Av = fam.param('A').default + 0.2 * rng.standard_normal(fam.param('A').size)
        bv = fam.param('b').default + 0.2 * rng.standard_normal(fam.param('b').size)
        for i, v in enumerate(Av):
            fn('cpg_update_A')(i, float(v))
        for i, v in enumerate(bv):
            fn('cpg_update_b')(i, float(v))
        fn('cpg_solve')()
        Px, Ax, (q, l, u) = canon_matrix_batches(fam, {'A': Av[None, :], 'b': bv[None, :]}, 1)
        ora = matrix_oracle_solve(fam, None, Ax, q, l, u)
        info = result.info.contents
        assert info.status == b'solved' and info.iter == int(ora['iter'][0])
        x = np.array([result.prim.contents.x[i] for i in range(2)])
        assert np.allclose(x, ora['x'][0, fam.variables[0].indices], rtol=1e-5, atol=1e-9)
        assert abs(info.obj_val - ora['obj'][0]) < 1e-8
