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


# ---------------------------------------------------------------------------------------------------------------------
# the conic plugin (IPMCUDAInterface, role of ECOSInterface) under the reference's writer
@pytest.mark.skipif(not refwriter.reference_available(), reason='reference tree not present')
def test_reference_writer_emits_the_conic_update_tree(tmp_path):
    """What the reference's emitters produce from IPMCUDAInterface's attribute values: ECOS's decision tree ('AbcGh' when anything of
    A, b, G is outdated, else 'c' / 'h' alone; cvxpygen/solvers/ecos.py:88-117) calling the shim, integer status, duals split into
    y and z, the settings table written through cpg_set_solver_<name>."""
    fam = families.portfolio_socp(20, 4, matrix_params=True)
    d = str(tmp_path / 'code')
    canon, iface, cfg = refwriter.write_reference_layout(fam, d, prefix='pf')
    solve_c = open(os.path.join(d, 'c', 'src', 'cpg_solve.c')).read()
    assert 'pf_cpg_b200_socp_shim_update(pf_Canon_Params.G->x, pf_Canon_Params.A->x, pf_Canon_Params.c, pf_Canon_Params.h, pf_Canon_Params.b);' in solve_c
    assert 'pf_cpg_b200_socp_shim_update(0, 0, pf_Canon_Params.c, 0, 0);' in solve_c
    assert 'pf_cpg_b200_socp_shim_solve();' in solve_c and 'pf_CPG_Info.status = pf_cpg_b200_socp_shim_info.status;' in solve_c
    assert 'pf_sol_z[' in solve_c and 'pf_sol_y[' in solve_c
    assert '(&pf_cpg_b200_socp_shim_settings)->maxit = maxit_new;' in solve_c
    ws_h = open(os.path.join(d, 'c', 'include', 'cpg_workspace.h')).read()
    assert 'extern CpgB200SocpShimInfo pf_cpg_b200_socp_shim_info;' in ws_h and '#include "cpg_b200_socp_shim.h"' in ws_h
    shim_h = open(os.path.join(d, 'c', 'solver_code', 'cpg_b200_socp_shim.h')).read()
    assert '#define CPG_SSHIM_HAS_G 1' in shim_h and '#define CPG_SSHIM_HAS_A 1' in shim_h
    # the emitted C compiles against the shim's declarations (no nvcc here)
    import subprocess
    for c in ('cpg_workspace.c', 'cpg_solve.c'):
        r = subprocess.run(['gcc', '-std=c99', '-fsyntax-only', '-I', os.path.join(d, 'c', 'include'), '-I', os.path.join(d, 'c', 'solver_code'),
                            os.path.join(d, 'c', 'src', c)], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-2000:]


@pytest.mark.gpu
@pytest.mark.parametrize('name,mat', [('refwriter_portfolio_socp_20_4', False), ('refwriter_portfolio_socp_mat_20_4', True)])
def test_pybind_conic_solve_equals_compiled_ecos(name, mat):
    """cpg_module.solve(upd, par) through the reference-emitted pybind module of the CONIC plugin: the emitted cpg_solve canonicalises
    on the host, hands c / b / h (and G / A values when F, d_sqrt are parameters) to the shim, the kernel solves a batch of one;
    compared with the compiled ECOS driven the reference's way (ECOS_updateData + ECOS_solve on the same canonical data).
    Runs in a fresh interpreter: every emitted module registers pybind classes of the same C++ names (cpg_params, ...), so a second
    cpg_module in one process shadows the first -- the reference has the same property (one generated module per process)."""
    import subprocess, sys
    from oracle import ref_ecos
    if not ref_ecos.available():
        pytest.skip('oracle/_ref/libecos_ref.so not built')
    d = _dir(name)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.abspath(__file__), d, str(int(mat))], capture_output=True, text=True,
                       env={**os.environ, 'PYTHONPATH': root + os.pathsep + os.path.join(root, 'tests')})
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]


def _conic_pybind_check(d, mat):
    from oracle import ref_ecos
    m = refwriter.load_module(d)
    fam = families.portfolio_socp(20, 4, matrix_params=mat)
    r = ref_ecos.RefECOS(fam.canon_data('c'), fam.canon_matrix('A'), fam.canon_data('b'), fam.canon_matrix('G'), fam.canon_data('h'),
                         fam.cone_dims['l'], fam.cone_dims['q'])
    rng = np.random.default_rng(8)
    m.set_solver_default_settings()
    for trial in range(4):
        vals = {'a': fam.param('a').default + 0.3 * rng.standard_normal(20), 'w_prev': np.full(20, 1 / 20)}
        if mat:
            vals['F'] = fam.param('F').default + 0.25 * rng.standard_normal(80)
            vals['d_sqrt'] = fam.param('d_sqrt').default * rng.uniform(0.5, 1.5, 20)
        pre = 'pf_' if mat else ''                     # classes carry the code-generation prefix, functions do not
        par = getattr(m, pre + 'cpg_params')(); upd = getattr(m, pre + 'cpg_updated')()
        for k, v in vals.items():
            if trial == 3 and k != 'a':
                continue                              # last trial: only `a` is marked outdated -> the 'c' branch of the tree
            setattr(par, k, list(np.asarray(v, dtype=float)))      # matrices: flat, column-major (what cpg_solver.py hands over)
            setattr(upd, k, True)
        if trial == 3:
            vals = {**prev, 'a': vals['a']}
        prev = vals
        res = m.solve(upd, par)
        th = fam.theta_default().copy()
        for k, v in vals.items():
            p_ = fam.param(k); th[p_.col:p_.col + p_.size] = v
        data = {k: np.asarray(fam.maps[k] @ th).ravel()[None, :] for k in ('c', 'b', 'h', 'A', 'G')}
        ora = r.solve_batch(c=data['c'], b=data['b'], h=data['h'], G=data['G'], A=data['A'])
        assert res.cpg_info.status == int(ora['exitflag'][0]) == 0 and res.cpg_info.iter == int(ora['iter'][0])
        for v in fam.variables:
            assert np.allclose(np.asarray(getattr(res.cpg_prim, v.name)), ora['x'][0, v.indices], rtol=1e-5, atol=1e-7), v.name
        for dv in fam.duals:
            assert np.allclose(np.asarray(getattr(res.cpg_dual, dv.name)).ravel(), ora[dv.vec][0, dv.indices], rtol=1e-4, atol=1e-6), dv.name
        assert abs(res.cpg_info.obj_val + ora['pcost'][0]) < 1e-7            # maximisation: cpg reports -(pcost)
    m.set_solver_maxit(3)
    assert m.solve(upd, par).cpg_info.iter == 3
    m.set_solver_default_settings()


if __name__ == '__main__':
    import sys
    _conic_pybind_check(sys.argv[1], bool(int(sys.argv[2])))
    print('ok')
