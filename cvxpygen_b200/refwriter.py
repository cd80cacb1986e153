"""Drive the REFERENCE's own code writer with the ADMM-CUDA plugin (boundary rows b1-b3, f3).

`cvxpygen.generator.Generator.generate` (cvxpygen/generator.py:65-95) runs: canonicalise -> `_setup_folder` ->
`solver_interface.generate_code` -> `CCodeWriter.write` -> compile.  Canonicalisation needs cvxpy, which this image lacks; every
later stage needs only numpy / scipy / jinja2.  This module runs those later stages with the reference's functions loaded from
the reference tree BY FILE PATH (nothing is copied):

    utils.write_workspace_prot / write_workspace_def     cvxpygen/utils.py:470-882    -> c/include/cpg_workspace.h, c/src/cpg_workspace.c
    utils.write_solve_prot / write_solve_def             cvxpygen/utils.py:885-1141   -> c/include/cpg_solve.h, c/src/cpg_solve.c
    templates/cpg_module.hpp.jinja2 + write_module_def   cvxpygen/utils.py:1163-1412  -> cpp/include/cpg_module.hpp, cpp/src/cpg_module.cpp

on a `Canon` bundle (the reference's own dataclasses, cvxpygen/mappings.py) filled from a `CanonFamily` the way
`canonicalizer.py:124-332` fills it, with `ADMMCUDAInterface` as the solver interface.  The emitted C / C++ is then compiled
(gcc / g++ with pybind11) and linked against the plugin's libcpg_b200.so: `cpg_module.solve(upd, par)` is the reference's
pybind entry, unmodified, running on the GPU.

One addition to the reference-emitted module (the NEW batched entry of SURVEY 8b): `cpp/src/cpg_module_batch.cpp`, emitted here,
defines `solve_batch(params: dict) -> dict`; it is registered by ONE line inserted into the emitted PYBIND11_MODULE body -- the
patch a maintainer would make to `write_module_def`.

The reference tree exists only in the build container; the generated directory (git-ignored, shipped to the GPU box like every
built library) is what the GPU tests load.
"""
import importlib.util
import io
import os
import shutil
import subprocess
import sys
import sysconfig
from types import SimpleNamespace

import numpy as np
import scipy.sparse as sp

REF_DEFAULT = os.environ.get('CPG_REFERENCE', '/root/reference')


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def reference_available(ref_root=REF_DEFAULT) -> bool:
    return os.path.exists(os.path.join(ref_root, 'cvxpygen', 'utils.py'))


def reference_modules(ref_root=REF_DEFAULT):
    """(utils, mappings) of the reference, loaded where they lie (they import numpy / scipy / jinja2 only)."""
    return (_load(os.path.join(ref_root, 'cvxpygen', 'utils.py'), 'cvxpygen_ref_utils'),
            _load(os.path.join(ref_root, 'cvxpygen', 'mappings.py'), 'cvxpygen_ref_mappings'))


def canon_from_family(fam, M):
    """A `Canon` (reference dataclasses) carrying the family, filled like cvxpygen/canonicalizer.py:124-332 does: user parameters
    in user-sparsity column order with `flat_usp` ending in 1.0, `p_id_to_mapping` CSR per canonical id, `p` holding the canonical
    data at the default parameters, the adjacency user parameter -> outdated canonical ids, duals as (vector name, indices)."""
    pinfo = M.ParameterInfo(
        col_to_name_usp={p.col: p.name for p in fam.params}, flat_usp=fam.theta_default(),
        id_to_col={i: p.col for i, p in enumerate(fam.params)}, ids=list(range(len(fam.params))),
        name_to_shape={p.name: p.shape for p in fam.params}, name_to_size_usp={p.name: p.size for p in fam.params},
        names=[p.name for p in fam.params], num=len(fam.params), writable={}, lower=None, upper=None)
    pc = M.ParameterCanon()
    pc.is_maximization = fam.is_maximization
    ids = ['P', 'q', 'd', 'A', 'l', 'u'] if fam.solver_type == 'quadratic' else ['c', 'd', 'A', 'b', 'G', 'h']
    for pid in ids:
        mp = fam.maps.get(pid)
        if mp is None:
            continue
        mp = sp.csr_matrix(mp)
        pc.p_id_to_mapping[pid] = mp
        pc.p_id_to_changes[pid] = bool(mp[:, :-1].nnz)
        pc.p_id_to_size[pid] = mp.shape[0]
        if pid in fam.patterns:
            pc.p[pid] = fam.canon_matrix(pid)
            pc.p_csc[pid] = pc.p[pid]
        else:
            pc.p[pid] = fam.canon_data(pid)
    pc.nonzero_d = bool(pc.p_id_to_mapping['d'].nnz) if 'd' in pc.p_id_to_mapping else False
    pc.quad_obj = 'P' in pc.p and pc.p['P'].nnz > 0
    pc.user_p_name_to_canon_outdated = {p.name: [pid for pid in pc.p_id_to_mapping if fam.changes(pid, [p.name])] for p in fam.params}
    n2o, n2i, n2s, n2sh, n2init, n2sym = {}, {}, {}, {}, {}, {}
    for v in fam.variables:
        n2o[v.name] = int(v.indices[0]); n2i[v.name] = np.asarray(v.indices); n2s[v.name] = len(v.indices); n2sh[v.name] = v.shape
        n2init[v.name] = np.zeros(v.shape) if len(v.indices) > 1 else 0.0
        n2sym[v.name] = False
    pv = M.PrimalVariableInfo(n2o, n2i, n2s, n2sh, n2init, n2sym, [False] * len(fam.variables))
    dv = M.DualVariableInfo({d.name: int(d.indices[0]) for d in fam.duals}, {d.name: (d.vec, np.asarray(d.indices)) for d in fam.duals},
                            {d.name: len(d.indices) for d in fam.duals}, {d.name: d.shape for d in fam.duals},
                            {d.name: (np.zeros(len(d.indices)) if len(d.indices) > 1 else 0.0) for d in fam.duals},
                            {d.name: d.vec for d in fam.duals})
    return M.Canon(pv, dv, pinfo, pc)


def _contiguous(idx):
    idx = np.asarray(idx)
    return len(idx) > 0 and np.array_equal(idx, idx[0] + np.arange(len(idx)))


# ---------------------------------------------------------------------------------------------------------------------
# the batched pybind entry (NEW): canonicalise the user-level rows on the host with the emitted canon_*_map
# tables, solve the batch through the shim (cpg_solve_batch_host), gather user-level variables / duals.
_BATCH_CPP = r'''/* Auto-generated by cvxpygen_b200: the batched entry of cpg_module (registered from PYBIND11_MODULE in cpg_module.cpp). */
#include <pybind11/pybind11.h>
#include <pybind11/numpy.h>
#include <pybind11/stl.h>
#include <stdexcept>
#include <string>
#include <vector>

extern "C" {
    #include "include/cpg_workspace.h"
    #include "include/cpg_solve.h"
}

namespace py = pybind11;
using arr = py::array_t<double, py::array::c_style | py::array::forcecast>;

namespace {
struct Par { const char* name; int col, size; };
struct Var { const char* name; int size; const int* idx; };
const Par PARAMS[] = { %(params)s };
%(idx_arrays)s
const Var PRIMS[] = { %(prims)s };
const Var DUALS[] = { %(duals)s };
constexpr int N_THETA = %(n_theta)d, N = %(n)d, M = %(m)d, NNZP = %(nnzP)d, NNZA = %(nnzA)d, ROW = %(row)d;

// rows [r0, r0 + nrows) of a canonical object = map (CSR, stored in the emitted cpg_csc as p/i/x) x theta
inline void apply_map(const cpg_csc* map, int nrows, const double* theta, double* out) {
  for (int r = 0; r < nrows; ++r) {
    double acc = 0.0;
    for (int k = map->p[r]; k < map->p[r + 1]; ++k) acc += map->x[k] * theta[map->i[k]];
    out[r] = acc;
  }
}
}  // namespace

py::dict %(p)ssolve_batch_cpp(py::dict params) {
  // ---- user-level rows: theta = defaults (cpg_params_vec), overwritten by what was passed (Fortran-order flattening is the caller's,
  //      like cpg_solve's get_param_value: cvxpygen/templates/cpg_solver.py.jinja2:26-34)
  py::ssize_t B = -1;
  std::vector<arr> given(sizeof(PARAMS) / sizeof(Par));
  std::vector<bool> has(given.size(), false);
  for (auto item : params) {
    const std::string key = py::cast<std::string>(item.first);
    bool found = false;
    for (size_t k = 0; k < given.size(); ++k)
      if (key == PARAMS[k].name) {
        arr a = py::cast<arr>(item.second);
        if (a.ndim() == 1 && PARAMS[k].size == 1) a = a.reshape({a.shape(0), (py::ssize_t)1});
        if (a.ndim() != 2 || a.shape(1) != PARAMS[k].size)
          throw std::invalid_argument("parameter " + key + ": expected an array of shape (B, " + std::to_string(PARAMS[k].size) + ")");
        if (B >= 0 && a.shape(0) != B) throw std::invalid_argument("inconsistent batch sizes");
        B = a.shape(0); given[k] = a; has[k] = true; found = true;
      }
    if (!found) throw py::attribute_error(key + " is not a parameter.");
  }
  if (B < 0) B = 1;
  std::vector<double> rows((size_t)B * ROW), theta(N_THETA);
  for (py::ssize_t b = 0; b < B; ++b) {
    for (int c = 0; c < N_THETA; ++c) theta[c] = %(p)scpg_params_vec[c];
    theta[N_THETA - 1] = 1.0;
    for (size_t k = 0; k < given.size(); ++k)
      if (has[k]) for (int e = 0; e < PARAMS[k].size; ++e) theta[PARAMS[k].col + e] = given[k].at(b, e);
    double* r = rows.data() + (size_t)b * ROW;
%(canon_rows)s
  }
  // ---- one batched launch through the C ABI (H2D, kernels, D2H)
  arr sol_x({B, (py::ssize_t)N}), sol_y({B, (py::ssize_t)(M > 0 ? M : 1)}), obj(B), pri(B), dua(B);
  py::array_t<int> iter(B), status(B);
  int rc;
  {
    py::gil_scoped_release nogil;
    rc = %(p)scpg_b200_shim_solve_batch((int)B, rows.data(), sol_x.mutable_data(), sol_y.mutable_data(), obj.mutable_data(),
                                      iter.mutable_data(), status.mutable_data(), pri.mutable_data(), dua.mutable_data());
  }
  if (rc != 0) throw std::runtime_error(std::string("cpg_b200 error: ") + CPG_B200_FN(cpg_b200_last_error)());
  // ---- retrieval (cpg_retrieve_prim / dual / info, cvxpygen/utils.py:950-985) for every instance
  py::dict prim, dual, info, out;
  for (const Var& v : PRIMS) {
    arr a({B, (py::ssize_t)v.size});
    for (py::ssize_t b = 0; b < B; ++b) for (int e = 0; e < v.size; ++e) a.mutable_at(b, e) = sol_x.at(b, v.idx[e]);
    prim[v.name] = a;
  }
  for (const Var& v : DUALS) {
    arr a({B, (py::ssize_t)v.size});
    for (py::ssize_t b = 0; b < B; ++b) for (int e = 0; e < v.size; ++e) a.mutable_at(b, e) = sol_y.at(b, v.idx[e]);
    dual[v.name] = a;
  }
  for (py::ssize_t b = 0; b < B; ++b) obj.mutable_at(b) = %(obj_sign)s(obj.at(b)%(plus_d)s);
  info["obj_val"] = obj; info["iter"] = iter; info["status"] = status; info["pri_res"] = pri; info["dua_res"] = dua;
  out["cpg_prim"] = prim; out["cpg_dual"] = dual; out["cpg_info"] = info; out["sol_x"] = sol_x; out["sol_y"] = sol_y;
  return out;
}

void %(p)scpg_b200_register_batch(py::module_& m) {
  m.def("solve_batch", &%(p)ssolve_batch_cpp, py::arg("params"),
        "Solve a batch of instances: params = {name: (B, size) float64}; one kernel launch through cpg_solve_batch_host.");
}
'''


def _batch_cpp(fam, canon, iface, prefix):
    pc = canon.parameter_canon
    n, m = iface.n_var, iface.n_eq + iface.n_ineq
    setup = iface.setup
    lines = []
    off = 0
    for pid, size in (('q', n), ('l', m), ('u', m)) + tuple((k, {'P': setup.nnzP, 'A': setup.nnzA}[k]) for k in ('P', 'A') if k in setup.mat_params):
        acc = '->x' if pid.isupper() else ''
        if pc.p_id_to_changes.get(pid):
            lines.append(f'    apply_map(&{prefix}canon_{pid}_map, {size}, theta.data(), r + {off});')
        else:
            lines.append(f'    for (int e = 0; e < {size}; ++e) r[{off} + e] = {prefix}Canon_Params.{pid}{acc}[e];')
        off += size
    idx_arrays, prims, duals = [], [], []
    for v in fam.variables:
        idx_arrays.append(f'const int IDX_P_{v.name}[] = {{' + ', '.join(str(int(i)) for i in v.indices) + '};')
        prims.append(f'{{"{v.name}", {len(v.indices)}, IDX_P_{v.name}}}')
    for d in fam.duals:
        idx_arrays.append(f'const int IDX_D_{d.name}[] = {{' + ', '.join(str(int(i)) for i in d.indices) + '};')
        duals.append(f'{{"{d.name}", {len(d.indices)}, IDX_D_{d.name}}}')
    ctx = dict(p=prefix, params=', '.join(f'{{"{p.name}", {p.col}, {p.size}}}' for p in fam.params),
               idx_arrays='\n'.join(idx_arrays), prims=', '.join(prims) or '{"", 0, nullptr}', duals=', '.join(duals) or '{"", 0, nullptr}',
               n_theta=fam.n_theta, n=n, m=m, nnzP=setup.nnzP, nnzA=setup.nnzA, row=setup.npb, canon_rows='\n'.join(lines),
               obj_sign='-' if pc.is_maximization else '', plus_d=f' + {prefix}Canon_Params.d' if pc.nonzero_d else '')
    return _BATCH_CPP % ctx


def write_reference_layout(fam, code_dir, prefix='', enable_settings=(), ref_root=REF_DEFAULT, gradient=False):
    """generator.py:65-95 minus canonicalisation and compilation, with the reference's own emitters.  Returns
    (canon, interface, configuration)."""
    from .solvers.admm_cuda import ADMMCUDAInterface
    from .solvers.ipm_cuda import IPMCUDAInterface
    conic = fam.solver_type != 'quadratic'
    if conic and gradient:
        raise ValueError('gradient code is emitted for the QP plugin (ADMM-CUDA)')
    U, M = reference_modules(ref_root)
    if prefix and not prefix[0].isalpha():       # generator.py:175-182
        prefix = f'_{prefix}'
    prefix = f'{prefix}_' if prefix else ''
    cfg = M.Configuration(code_dir, 'IPM-CUDA' if conic else 'ADMM-CUDA', prefix, bool(gradient), False, 0)
    canon = canon_from_family(fam, M)
    iface = (IPMCUDAInterface if conic else ADMMCUDAInterface)(family=fam, enable_settings=enable_settings)
    # _setup_folder (generator.py:99-108)
    shutil.rmtree(code_dir, ignore_errors=True)
    for sub in ('c/src', 'c/include', 'c/build', 'cpp/src', 'cpp/include'):
        os.makedirs(os.path.join(code_dir, sub))
    solver_code_dir = os.path.join(code_dir, 'c', 'solver_code')
    # _run_solver_code_generation (generator.py:124-146)
    if conic:
        iface.generate_reference_layout_code(code_dir, solver_code_dir, canon, prefix)
    else:
        iface.generate_code(cfg, code_dir, solver_code_dir, os.path.join(ref_root, 'cvxpygen'), canon, gradient, prefix)
    # CCodeWriter._write_workspace / _write_solve / _write_python_module (writer.py:96-139, 596-609)
    pvi, dvi, pi, pc = canon.prim_variable_info, canon.dual_variable_info, canon.parameter_info, canon.parameter_canon
    inc, src = os.path.join(code_dir, 'c', 'include'), os.path.join(code_dir, 'c', 'src')
    U.write_file(os.path.join(inc, 'cpg_workspace.h'), 'w', U.write_workspace_prot, cfg, pvi, dvi, pi, pc, iface, True)
    U.write_file(os.path.join(src, 'cpg_workspace.c'), 'w', U.write_workspace_def, cfg, pvi, dvi, pi, pc, iface, True)
    U.write_file(os.path.join(inc, 'cpg_solve.h'), 'w', U.write_solve_prot, cfg, pvi, dvi, pi, pc, iface, None)
    U.write_file(os.path.join(src, 'cpg_solve.c'), 'w', U.write_solve_def, cfg, pvi, dvi, pi, pc, iface, None)
    U.render_template_to_file('cpg_module.hpp.jinja2', os.path.join(code_dir, 'cpp', 'include'),
                              U.module_hpp_context(cfg, pi, pvi, dvi, iface, iface))
    buf = io.StringIO()
    U.write_module_def(buf, cfg, pvi, dvi, pi, iface, iface)
    txt = buf.getvalue()
    if not conic:
        # the one-line patch to write_module_def a maintainer would make for the batched entry
        txt = txt.replace('namespace py = pybind11;\n', f'namespace py = pybind11;\nvoid {prefix}cpg_b200_register_batch(py::module_& m);\n', 1)
        k = txt.rindex('\n}')
        txt = txt[:k] + f'\n    {prefix}cpg_b200_register_batch(m);\n' + txt[k:]
    with open(os.path.join(code_dir, 'cpp', 'src', 'cpg_module.cpp'), 'w') as f:
        f.write(txt)
    if not conic:
        with open(os.path.join(code_dir, 'cpp', 'src', 'cpg_module_batch.cpp'), 'w') as f:
            f.write(_batch_cpp(fam, canon, iface, prefix))
    with open(os.path.join(code_dir, '__init__.py'), 'w') as f:
        f.write('')
    return canon, iface, cfg


def compile_reference_layout(code_dir, verbose=False):
    """Role of the reference's build (cmake -> libcpg.a, then setup.py build_ext; cvxpygen/compiler.py:24-31,
    templates/setup.py.jinja2:74-117) without cmake: nvcc for the CUDA library, gcc for the emitted C, g++ + pybind11 for the
    emitted module; everything in-tree.  Returns the path of the built cpg_module extension."""
    from . import codegen, codegen_ipm
    import pybind11
    sol = os.path.join(code_dir, 'c', 'solver_code')
    conic = os.path.exists(os.path.join(sol, 'cpg_b200_socp_shim.c'))
    if conic:
        lib = codegen_ipm.compile_ipm_solver_sources(sol, os.path.join(code_dir, 'libcpg_b200.so'), verbose=verbose)
    else:
        lib = codegen.compile_solver_sources(sol, os.path.join(code_dir, 'libcpg_b200.so'), verbose=verbose)
    inc = os.path.join(code_dir, 'c', 'include')
    objs = []
    for c in (os.path.join(code_dir, 'c', 'src', 'cpg_workspace.c'), os.path.join(code_dir, 'c', 'src', 'cpg_solve.c'),
              os.path.join(sol, 'cpg_b200_socp_shim.c' if conic else 'cpg_b200_shim.c')):
        o = os.path.join(code_dir, 'c', 'build', os.path.basename(c)[:-2] + '.o')
        cmd = ['gcc', '-O2', '-fPIC', '-std=c99', '-I', inc, '-I', sol, '-c', c, '-o', o]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode:
            raise RuntimeError(f'gcc failed on the emitted {os.path.basename(c)}:\n' + res.stderr[-4000:])
        objs.append(o)
    ext = os.path.join(code_dir, 'cpg_module' + sysconfig.get_config_var('EXT_SUFFIX'))
    cmd = ['g++', '-O2', '-fPIC', '-shared', '-std=c++17', '-fvisibility=hidden',
           '-I', pybind11.get_include(), '-I', sysconfig.get_paths()['include'], '-I', os.path.join(code_dir, 'cpp', 'include'),
           '-I', os.path.join(code_dir, 'c'), '-I', inc, '-I', sol,
           os.path.join(code_dir, 'cpp', 'src', 'cpg_module.cpp')] + \
          ([] if conic else [os.path.join(code_dir, 'cpp', 'src', 'cpg_module_batch.cpp')]) + objs + \
          [lib, '-Wl,-rpath,$ORIGIN', '-o', ext]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode:
        raise RuntimeError('g++ failed on the emitted cpg_module.cpp:\n' + res.stderr[-4000:])
    if verbose:
        sys.stdout.write(res.stderr)
    return ext


def load_module(code_dir):
    """import the built pybind extension of a reference-layout directory (a fresh module object per directory)."""
    ext = os.path.join(code_dir, 'cpg_module' + sysconfig.get_config_var('EXT_SUFFIX'))
    if not os.path.exists(ext):
        raise RuntimeError(f'{ext} is missing: run refwriter.compile_reference_layout where the reference tree exists')
    spec = importlib.util.spec_from_file_location('cpg_module', ext)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


# the reference-layout directories built ahead of time (they travel to the GPU box): name -> (family builder, prefix)
def standard_layouts():
    from . import families
    return {
        'refwriter_mpc_6_3_10': (lambda: families.mpc(6, 3, 10), ''),                   # vectors only (l, u change): the 'lu' branch
        'refwriter_nonneg_LS_3_2_A': (lambda: families.nonneg_ls(3, 2, name='nonneg_LS_3_2_A'), 'nnls'),   # A + l/u change; prefix
        # the conic plugin (IPM-CUDA, role of ECOSInterface): c / b change ('AbcGh' and 'c' branches), and F / d_sqrt in A / G
        'refwriter_portfolio_socp_20_4': (lambda: families.portfolio_socp(20, 4), ''),
        'refwriter_portfolio_socp_mat_20_4': (lambda: families.portfolio_socp(20, 4, matrix_params=True), 'pf'),
    }


def build_standard_layouts(force=False, ref_root=REF_DEFAULT):
    from .standard import GENERATED_DIR
    out = {}
    for name, (fn, prefix) in standard_layouts().items():
        d = os.path.join(GENERATED_DIR, name)
        ext = os.path.join(d, 'cpg_module' + sysconfig.get_config_var('EXT_SUFFIX'))
        if force or not os.path.exists(ext):
            if not reference_available(ref_root):
                continue                      # the GPU box: use what was built in the container
            write_reference_layout(fn(), d, prefix=prefix, ref_root=ref_root)
            compile_reference_layout(d)
        out[name] = d
    return out
