"""Host-side checks that need no GPU: code generation, the nvcc build for sm_100a, the C-ABI surface
(every symbol include/cpg_b200.h declares is exported, plus the reference-compatible per-family symbols),
error behaviour of the front end, and loud failure without a CUDA device."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from cvxpygen_b200 import cpg, families, runtime, standard

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def ls_dir(tmp_path_factory):
    d = str(tmp_path_factory.mktemp('gen') / 'nonneg_LS')
    cpg.generate_code(families.nonneg_ls(3, 2), code_dir=d, solver='ADMM-CUDA', batch_params=['b'], prefix='t1')
    return d


def declared_functions():
    txt = open(os.path.join(ROOT, 'include', 'cpg_b200.h')).read()
    return sorted(set(re.findall(r'CPG_B200_FN\((\w+)\)\s*\(', txt)))


def test_generated_tree_mirrors_reference_layout(ls_dir):
    for rel in ('cpg_solver.py', 'cpg_module.py', 'cpg_meta.json', '__init__.py', 'libcpg_b200.so',
                'c/include/cpg_b200.h', 'c/include/cpg_family.h', 'c/include/cpg_blob_layout.h',
                'c/include/cpg_workspace.h', 'c/include/cpg_solve.h', 'c/src/cpg_blob.c', 'c/src/cpg_solve.c',
                'c/solver_code/admm_kernel.cuh', 'c/solver_code/cpg_b200_module.cu'):
        assert os.path.exists(os.path.join(ls_dir, rel)), rel


def test_library_exports_every_declared_symbol(ls_dir):
    lib = C.CDLL(os.path.join(ls_dir, 'libcpg_b200.so'))
    fns = declared_functions()
    assert {'cpg_b200_init', 'cpg_solve_batch_device', 'cpg_solve_batch_host', 'cpg_b200_dims',
            'cpg_gradient_batch_device', 'cpg_gradient_batch_host'} <= set(fns)
    for fn in fns:
        assert hasattr(lib, 't1_' + fn), fn
    # reference-compatible per-family interface (cvxpygen/utils.py:1087-1141)
    for fn in ('cpg_update_A', 'cpg_update_b', 'cpg_solve', 'cpg_retrieve_prim', 'cpg_retrieve_dual', 'cpg_retrieve_info',
               'cpg_set_solver_default_settings', 'cpg_set_solver_max_iter', 'cpg_set_solver_eps_abs',
               'cpg_set_solver_warm_starting', 'CPG_Result', 'CPG_Prim', 'CPG_Dual', 'CPG_Info', 'cpg_params_vec'):
        assert hasattr(lib, 't1_' + fn), fn


def test_sass_is_sm100a_with_tma_bulk_copy(ls_dir):
    out = subprocess.run(['cuobjdump', '-sass', os.path.join(ls_dir, 'libcpg_b200.so')], capture_output=True, text=True).stdout
    assert 'sm_100a' in out
    assert 'UBLKCP' in out            # cp.async.bulk global->shared (TMA) staging of the constants blob
    assert 'DFMA' in out              # fp64 arithmetic


def test_dims_settings_without_gpu_and_loud_init_failure(ls_dir):
    mod = runtime.Module(ls_dir)
    assert (mod.dims.n_var, mod.dims.n_con, mod.dims.n_param, mod.dims.n_prim, mod.dims.n_dual) == (5, 5, 3, 2, 2)
    s = mod.settings
    assert (s.max_iter, s.check_termination, s.adaptive_rho, s.warm_start) == (4000, 25, 1, 0)
    assert (s.eps_abs, s.eps_rel, s.eps_prim_inf, s.eps_dual_inf, s.alpha) == (1e-3, 1e-3, 1e-4, 1e-4, 1.6)
    with pytest.raises(AttributeError):
        mod.set_solver_setting('polish', 1)          # not enabled, like the reference's disabled settings
    mod.set_solver_max_iter(17)
    assert mod.settings.max_iter == 17
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if not has_gpu:
        with pytest.raises(RuntimeError):            # no CPU fallback: the product path fails loudly
            mod.solve_batch({'b': np.zeros((2, 3))})


def test_missing_library_fails_loudly(tmp_path):
    d = str(tmp_path / 'nolib')
    cpg.generate_code(families.nonneg_ls(3, 2), code_dir=d, batch_params=['b'], wrapper=False)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        runtime.Module(d)


def test_pack_params_shapes_and_errors(ls_dir):
    mod = runtime.Module(ls_dir)
    P = mod.pack_params({'b': np.arange(12.0).reshape(4, 3)})
    assert P.shape == (4, 3) and P[2, 1] == 7.0
    assert mod.pack_params({'b': np.array([1.0, 2.0, 3.0])}).shape == (1, 3)
    assert np.array_equal(mod.pack_params({}), np.asarray(mod.meta['params'][1]['default'])[None])
    with pytest.raises(AttributeError, match='is not a parameter'):
        mod.pack_params({'nope': np.zeros(3)})
    with pytest.raises(ValueError, match='shared'):
        mod.pack_params({'A': np.zeros((2, 3))})


def test_matrix_parameter_is_flattened_in_fortran_order(tmp_path):
    """A (B, *shape) batched parameter is flattened column-major per instance (TPL/cpg_solver.py.jinja2:26-34)."""
    fam = families.mpc(2, 1, 2)
    d = str(tmp_path / 'm')
    cpg.generate_code(fam, code_dir=d, batch_params=['x_init'], wrapper=False)
    import json
    meta = json.load(open(os.path.join(d, 'cpg_meta.json')))
    assert [v['name'] for v in meta['variables']] == ['U', 'X'] and meta['variables'][1]['shape'] == [2, 3]
    assert meta['duals'][0]['shape'] == [2, 2] and meta['params'][0]['batched']


def test_generate_code_argument_errors():
    with pytest.raises(ValueError, match='Unsupported solver'):
        cpg.generate_code(families.nonneg_ls(), solver='GUROBI', wrapper=False)


def test_standard_families_are_built_or_buildable():
    d = standard.build('nonneg_LS_3_2')
    assert os.path.exists(os.path.join(d, 'libcpg_b200.so'))


def test_family_from_canonical_qp_data():
    """CanonFamily.from_canonical_qp: OSQP's own basic_qp (osqp_sources/tests/basic_qp/generate_problem.py) as a family whose
    parameters are its canonical vectors -- and, with matrix_params, its matrix entries; the offline setup accepts both."""
    import scipy.sparse as sp
    from cvxpygen_b200.ir import CanonFamily
    from cvxpygen_b200.offline.qp_setup import setup_qp_family
    from oracle.admm_numpy import AdmmOracle
    P = sp.csc_matrix([[4.0, 1.0], [1.0, 2.0]]); q = np.array([1.0, 1.0])
    A = sp.csc_matrix([[1.0, 1.0], [1.0, 0.0], [0.0, 1.0], [0.0, 1.0]])
    l = np.array([1.0, 0.0, 0.0, -np.inf]); u = np.array([1.0, 0.7, 0.7, np.inf])
    fam = CanonFamily.from_canonical_qp('basic_qp', P, q, A, np.clip(l, -1e30, 1e30), np.clip(u, -1e30, 1e30))
    assert [p.name for p in fam.params] == ['q', 'l', 'u'] and fam.n_eq == 1
    assert np.array_equal(fam.canon_data('q'), q) and fam.canon_matrix('P').nnz == 3
    st = setup_qp_family(fam, ['q', 'l', 'u'])
    assert st.npb == 2 + 4 + 4 and not st.mat_params
    fam2 = CanonFamily.from_canonical_qp('basic_qp_m', P, q, A, np.clip(l, -1e30, 1e30), np.clip(u, -1e30, 1e30), matrix_params=True)
    st2 = setup_qp_family(fam2, ['q', 'l', 'u', 'P', 'A'])
    assert st2.mat_params == ['P', 'A'] and len(st2.mat_blob) > 0
    sol = AdmmOracle(fam.canon_matrix('P'), q, fam.canon_matrix('A'), fam.canon_data('l'), fam.canon_data('u'),
                     eps_abs=1e-9, eps_rel=1e-9).solve_batch(B=1)
    assert np.allclose(sol['x'][0], [0.3, 0.7], atol=1e-6) and abs(sol['obj'][0] - 1.88) < 1e-6      # OSQP's known answer
