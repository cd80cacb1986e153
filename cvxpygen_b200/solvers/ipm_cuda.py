"""`IPMCUDAInterface` -- the solver-plugin class of the IPM-CUDA backend (boundary b1, conic families).

The conic twin of `ADMMCUDAInterface`: the attribute set the reference's writer reads from a `SolverInterface`
(reference: cvxpygen/solvers/_interface.py:82-258), with the values of the plugin it stands beside --
`ECOSInterface` (cvxpygen/solvers/ecos.py:16-134): conic canonical form (c, d, A, b, G, h) + cone dimensions, integer exit
flags, duals split into y (equalities) and z (cones), ECOS's settings table.  It carries no solver arithmetic:
`generate_code` runs the offline setup (offline/socp_setup.py) and emits the CUDA sources (codegen_ipm.py).
"""
import os

from .admm_cuda import Setting, WorkspacePointerInfo, UpdatePendingLogic, ParameterUpdateLogic

try:                                              # reference present: be a real plugin
    from cvxpygen.solvers import SolverInterface as _RefBase   # pragma: no cover
    _BASES = (_RefBase,)
except Exception:                                 # cvxpy / cvxpygen absent
    _BASES = (object,)


def _upd(G, A, c, h, b):
    """The emitted call into the shim: canonical arrays that are (possibly) outdated, 0 for the others."""
    mat = lambda on, name: f'{{prefix}}Canon_Params.{name}->x' if on else '0'
    vec = lambda on, name: f'{{prefix}}Canon_Params.{name}' if on else '0'
    return f'{{prefix}}cpg_b200_socp_shim_update({mat(G, "G")}, {mat(A, "A")}, {vec(c, "c")}, {vec(h, "h")}, {vec(b, "b")})'


_STGS = [('feastol', 'cpg_float', '1e-8', None), ('abstol', 'cpg_float', '1e-8', None), ('reltol', 'cpg_float', '1e-8', None),
         ('feastol_inacc', 'cpg_float', '1e-4', None), ('abstol_inacc', 'cpg_float', '5e-5', None),
         ('reltol_inacc', 'cpg_float', '5e-5', None), ('maxit', 'cpg_int', '100', 'max_iters')]      # cvxpygen/solvers/ecos.py:59-67


class IPMCUDAInterface(*_BASES):
    solver_name = 'IPM-CUDA'
    cvxpy_solver_name = 'ECOS'          # canonicalise through cvxpy's ECOS (conic) path
    solver_type = 'conic'
    supports_quad_obj = False
    canon_p_ids = ['c', 'd', 'A', 'b', 'G', 'h']
    canon_p_ids_constr_vec = ['b', 'h']
    dual_var_split = True
    dual_var_names = ['y', 'z']
    # same decision tree as ECOS's (cvxpygen/solvers/ecos.py:88-117): anything of A, b, G outdated -> everything is handed over (the
    # kernel re-equilibrates per instance, like ECOS_updateData); only c / only h -> that vector.  The 'init' branch of ECOS
    # (ECOS_setup on the first call) has no counterpart: the shim initialises the device library on its first solve.
    parameter_update_structure = {
        'AbcGh': ParameterUpdateLogic(UpdatePendingLogic(['A', 'b', 'G'], '||', ['c', 'h']), _upd(1, 1, 1, 1, 1)),
        'c': ParameterUpdateLogic(UpdatePendingLogic(['c']), _upd(0, 0, 1, 0, 0)),
        'h': ParameterUpdateLogic(UpdatePendingLogic(['h']), _upd(0, 0, 0, 1, 0)),
    }
    solve_function_call = '{prefix}cpg_b200_socp_shim_solve()'
    header_files = ['"cpg_b200_socp_shim.h"']
    cmake_headers = ['${CMAKE_CURRENT_SOURCE_DIR}/*.h', '${CMAKE_CURRENT_SOURCE_DIR}/*.cuh']
    cmake_sources = ['${CMAKE_CURRENT_SOURCE_DIR}/cpg_b200_socp_shim.c']
    inmemory_preconditioning = False    # equilibration happens at generation time (shared matrices) or inside the kernel (per instance)
    # workspace: the shim's result block lives in cpg_workspace.{h,c} (declare_workspace / define_workspace), prefixed
    ws_statically_allocated_in_solver_code = False
    ws_ptrs = WorkspacePointerInfo(
        objective_value='cpg_b200_socp_shim_info.pcost',
        iterations='cpg_b200_socp_shim_info.iter',
        status='cpg_b200_socp_shim_info.status',
        primal_residual='cpg_b200_socp_shim_info.pres',
        dual_residual='cpg_b200_socp_shim_info.dres',
        primal_solution='sol_x',
        dual_solution='sol_{dual_var_name}')
    sol_statically_allocated = True     # CPG_Prim / CPG_Dual point into {prefix}sol_x / sol_y / sol_z
    status_is_int = True                # ECOS exit flags: 0 optimal, 1 / 2 primal / dual infeasible, +10 inaccurate, -1 maxit, ...
    numeric_types = {'float': 'double', 'int': 'int'}
    stgs_dynamically_allocated = False
    stgs_requires_extra_struct_type = False
    stgs_direct_write_ptr = '(&{prefix}cpg_b200_socp_shim_settings)'
    stgs_reset_function = {'name': 'cpg_b200_socp_shim_default_settings', 'ptr': '&{prefix}cpg_b200_socp_shim_settings'}
    stgs = {n: Setting(t, d, True, cv) for n, t, d, cv in _STGS}
    docu = 'DESIGN.md'

    def __init__(self, data=None, p_prob=None, enable_settings=(), family=None):
        if family is None:                        # reference-style construction (needs cvxpygen + cvxpy)
            cd = p_prob.cone_dims
            self.check_unsupported_cones(cd)
            canon_constants = {'n': p_prob.x.size, 'm': data['G'].shape[0], 'p': cd.zero, 'l': cd.nonneg,
                               'n_cones': len(cd.soc), 'q': list(cd.soc), 'e': cd.exp}
            super().__init__(self.solver_name, p_prob.x.size, cd.zero, data['G'].shape[0], p_prob, canon_constants,
                             list(enable_settings))
        else:                                     # cvxpy-free construction
            self.n_var, self.n_eq, self.n_ineq = family.n_var, family.n_eq, family.n_ineq
            self.canon_constants = {'n': family.n_var, 'm': family.n_ineq, 'p': family.n_eq, 'l': family.cone_dims.get('l', 0),
                                    'n_cones': len(family.cone_dims.get('q', [])), 'q': list(family.cone_dims.get('q', [])), 'e': 0}
            self.enable_settings = list(enable_settings)
        self.family = family
        self.setup = None
        has_eq = self.n_eq > 0          # like ECOSInterface: A and b are passed as 0 when there are no equalities (ecos.py:96-101)
        self.parameter_update_structure = {
            'AbcGh': ParameterUpdateLogic(UpdatePendingLogic(['A', 'b', 'G'] if has_eq else ['G'], '||', ['c', 'h']),
                                          _upd(1, has_eq, 1, 1, has_eq)),
            'c': ParameterUpdateLogic(UpdatePendingLogic(['c']), _upd(0, 0, 1, 0, 0)),
            'h': ParameterUpdateLogic(UpdatePendingLogic(['h']), _upd(0, 0, 0, 1, 0)),
        }
        self.stgs = {n: Setting(t, d, True, cv) for n, t, d, cv in _STGS}
        if self.enable_settings and family is not None:
            for n, st_ in self.stgs.items():
                st_.enabled = n in self.enable_settings

    @property
    def stgs_names_enabled(self):
        return [n for n, s in self.stgs.items() if s.enabled]

    @property
    def stgs_names_to_type(self):
        return {n: s.type for n, s in self.stgs.items() if s.enabled}

    @property
    def stgs_names_to_default(self):
        return {n: s.default for n, s in self.stgs.items() if s.enabled}

    @property
    def stgs_translation(self):
        return {s.name_cvxpy: n for n, s in self.stgs.items() if s.enabled and s.name_cvxpy is not None}

    @staticmethod
    def check_unsupported_cones(cone_dims) -> None:
        """LP and second-order cones only (the reference's ECOS plugin rejects exponential cones the same way,
        cvxpygen/solvers/ecos.py:121-125)."""
        if getattr(cone_dims, 'exp', 0) > 0:
            raise ValueError('Code generation with IPM-CUDA and exponential cones is not supported yet.')
        if len(getattr(cone_dims, 'psd', []) or []) > 0 or len(getattr(cone_dims, 'p3d', []) or []) > 0:
            raise ValueError('Code generation with IPM-CUDA supports the nonnegative and second-order cones only.')

    @staticmethod
    def ret_prim_func_exists(variable_info) -> bool:
        return True

    @staticmethod
    def ret_dual_func_exists(dual_variable_info) -> bool:
        return True

    # ---- build-system hooks (cvxpygen/solvers/_interface.py:203-236)
    def cmake_context_extra(self) -> dict:
        sdir = '${CMAKE_CURRENT_SOURCE_DIR}/solver_code'
        return {'solver_code_cmake_include_dir': sdir, 'extra_cmake_include_dirs': [sdir], 'packages': ['CUDAToolkit'],
                'cmake_target_link_libs': ['${CMAKE_CURRENT_SOURCE_DIR}/../libcpg_b200.so', 'CUDA::cudart'], 'cmake_definitions': []}

    def setup_py_context(self) -> dict:
        return {'solver_code_include_dir': "os.path.join('c', 'solver_code')", 'extra_solver_include_dirs': [],
                'extra_cpp_include_dirs': ["os.path.join('c', 'solver_code')"], 'extra_lib_names_windows': None,
                'extra_lib_names_unix': ['cpg_b200'], 'extra_objects': ["os.path.join('libcpg_b200.so')"], 'license': 'Apache 2.0'}

    # ---- workspace hooks (cvxpygen/utils.py:661-662, 862-863)
    def declare_workspace(self, f, prefix, parameter_canon) -> None:
        f.write('\n// IPM-CUDA workspace: canonical solution, solver info and settings of the last cpg_solve()\n')
        f.write(f'extern cpg_float {prefix}sol_x[{self.n_var}];\n')
        f.write(f'extern cpg_float {prefix}sol_y[{max(self.n_eq, 1)}];\n')
        f.write(f'extern cpg_float {prefix}sol_z[{max(self.n_ineq, 1)}];\n')
        f.write(f'extern CpgB200SocpShimInfo {prefix}cpg_b200_socp_shim_info;\n')
        f.write(f'extern CpgB200SocpSettings {prefix}cpg_b200_socp_shim_settings;\n')

    def define_workspace(self, f, prefix, parameter_canon) -> None:
        f.write('\n// IPM-CUDA workspace\n')
        f.write(f'cpg_float {prefix}sol_x[{self.n_var}];\n')
        f.write(f'cpg_float {prefix}sol_y[{max(self.n_eq, 1)}];\n')
        f.write(f'cpg_float {prefix}sol_z[{max(self.n_ineq, 1)}];\n')
        f.write(f'CpgB200SocpShimInfo {prefix}cpg_b200_socp_shim_info = {{0, 0, -7, 0, 0}};\n')
        f.write(f'CpgB200SocpSettings {prefix}cpg_b200_socp_shim_settings = {{100, 0, 1e-8, 1e-8, 1e-8, 1e-4, 5e-5, 5e-5}};\n')

    def generate_reference_layout_code(self, code_dir, solver_code_dir, canon, prefix, compile=False, threads=None):
        """What `generate_code` does when the REFERENCE's generator drives the plugin (cvxpygen/generator.py:124-146): the CUDA sources
        of the canonical-level family (user parameters = the canonical arrays the emitted cpg_solve hands over; a matrix is one of
        them iff a user parameter enters it) + the shim, flat in <code_dir>/c/solver_code."""
        import numpy as np
        from .. import codegen_ipm
        from ..ir import CanonFamily
        from ..offline.socp_setup import setup_socp_family, DEFAULT_THREADS
        from ..shim_socp import write_socp_shim
        pc = canon.parameter_canon
        changes = {k: bool(v) for k, v in pc.p_id_to_changes.items()}
        mats = tuple(k for k in ('G', 'A') if changes.get(k) and (k != 'A' or self.n_eq > 0))
        name = getattr(self.family, 'name', None) or os.path.basename(os.path.abspath(code_dir))
        cc = self.canon_constants
        A = pc.p['A'] if self.n_eq else __import__('scipy.sparse').sparse.csc_matrix((0, self.n_var))
        b = pc.p['b'] if self.n_eq else np.zeros(0)
        fam = CanonFamily.from_canonical_conic(name + '_canonical', pc.p['c'], A, b, pc.p['G'], pc.p['h'], cc['l'], cc['q'],
                                               matrix_params=mats)
        fam.is_maximization = False      # the sign flip is applied by the emitted cpg_retrieve_info (cvxpygen/utils.py:980)
        batch = [p.name for p in fam.params]
        self.setup = setup_socp_family(fam, batch, threads=int(threads or DEFAULT_THREADS))
        cprefix = prefix or ''
        codegen_ipm.write_ipm_solver_sources(self.setup, solver_code_dir, prefix=cprefix, threads=threads)
        write_socp_shim(self.setup, solver_code_dir, cprefix, matrices=mats)
        if compile:
            codegen_ipm.compile_ipm_solver_sources(solver_code_dir, os.path.join(code_dir, 'libcpg_b200.so'))
        return self.setup

    def generate_code(self, configuration, code_dir, solver_code_dir, cvxpygen_directory, canon, gradient, prefix,
                      batch_params=None, compile=True, threads=None):
        from .. import codegen_ipm
        from ..ir import CanonFamily
        from ..offline.socp_setup import setup_socp_family, DEFAULT_THREADS
        if gradient:
            raise ValueError('gradient=True is generated for the QP path (ADMM-CUDA) only')
        fam = self.family if self.family is not None else CanonFamily.from_reference_canon(
            getattr(configuration, 'code_dir', 'problem'), canon, self)
        setup = setup_socp_family(fam, batch_params, threads=int(threads or DEFAULT_THREADS))
        codegen_ipm.write_ipm_code(setup, code_dir, prefix=(prefix or '').rstrip('_'), threads=threads)
        if compile:
            codegen_ipm.compile_ipm_code(code_dir)
        return setup
