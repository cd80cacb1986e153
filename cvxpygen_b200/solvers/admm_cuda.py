"""`ADMMCUDAInterface` -- the solver-plugin class of the ADMM-CUDA backend (boundary b1).

It carries every attribute and hook the reference's writer reads from a `SolverInterface`
(reference: cvxpygen/solvers/_interface.py:82-258; the OSQP instance of it is cvxpygen/solvers/osqp.py:16-163) with values
that make the reference's OWN emitters -- `write_workspace_def/prot`, `write_solve_def/prot`, `write_module_def`
(cvxpygen/utils.py:470-1141, 1163-1412) -- produce C / C++ that compiles and links against the CUDA library this plugin
generates (exercised by cvxpygen_b200/refwriter.py and tests/test_refwriter.py, which drive those emitters where they lie
under /root/reference):

  * `generate_code(...)` writes the CUDA sources of a *canonical-level* family into <code_dir>/c/solver_code (the canonical
    vectors q, l, u -- and the entries of P / A when a user parameter enters them -- are the per-instance inputs of the
    kernels) together with `cpg_b200_shim.{h,c}`, the glue the emitted `cpg_solve()` calls;
  * `parameter_update_structure` mirrors OSQP's table (cvxpygen/solvers/osqp.py:20-61): the same decision tree, with
    `{prefix}cpg_b200_shim_update_mat / _vec` instead of `osqp_update_data_mat / _vec`;
  * `solve_function_call` = `{prefix}cpg_b200_shim_solve()` (one instance = a batch of one on the GPU, no CPU fallback);
  * `ws_ptrs` point at the shim's workspace, which `declare_workspace` / `define_workspace` emit into cpg_workspace.{h,c}
    with the code-generation prefix (ws_statically_allocated_in_solver_code = False, like ECOS: cvxpygen/solvers/ecos.py:32);
  * settings are written straight into `{prefix}cpg_b200_shim_settings` (stgs_direct_write_ptr), reset by
    `cpg_b200_shim_default_settings`.

When the reference package is importable the class derives from its `QPCanonMixin` + `SolverInterface`; otherwise (this
container: cvxpy is absent) it stands alone and is constructed from a `CanonFamily`.  The class carries no solver arithmetic.
"""
import os
from dataclasses import dataclass
from typing import List

try:                                              # reference present: be a real plugin
    from cvxpygen.solvers import SolverInterface as _RefBase, QPCanonMixin as _RefMixin   # pragma: no cover
    from cvxpygen.mappings import WorkspacePointerInfo, UpdatePendingLogic, ParameterUpdateLogic, Setting  # pragma: no cover
    _BASES = (_RefMixin, _RefBase)
except Exception:                                 # cvxpy / cvxpygen absent: same-shaped stand-ins (cvxpygen/mappings.py:94-145)
    _BASES = (object,)

    @dataclass
    class Setting:
        type: str
        default: str
        enabled: bool = True
        name_cvxpy: str = None

    @dataclass
    class WorkspacePointerInfo:
        objective_value: str
        iterations: str
        status: str
        primal_residual: str
        dual_residual: str
        primal_solution: str
        dual_solution: str
        settings: str = None

    @dataclass
    class UpdatePendingLogic:
        parameters_outdated: List[str]
        operator: str = None
        functions_if_false: List[str] = None
        extra_condition: str = None
        extra_condition_operator: str = None

    @dataclass
    class ParameterUpdateLogic:
        update_pending_logic: UpdatePendingLogic
        function_call: str


def _vec_call(q, l, u):
    arg = lambda on, name: f'{{prefix}}Canon_Params.{name}' if on else '0'
    return f'{{prefix}}cpg_b200_shim_update_vec({arg(q, "q")}, {arg(l, "l")}, {arg(u, "u")})'


def _mat_call(P, A):
    arg = lambda on, name: f'{{prefix}}Canon_Params.{name}->x' if on else '0'
    return f'{{prefix}}cpg_b200_shim_update_mat({arg(P, "P")}, {arg(A, "A")})'


# the twelve settings of cvxpygen/solvers/osqp.py:102-115, in that order (= field order of CpgB200ShimSettings)
_STGS = [('max_iter', 'cpg_int', '4000', True, None), ('eps_abs', 'cpg_float', '1e-3', True, None),
         ('eps_rel', 'cpg_float', '1e-3', True, None), ('eps_prim_inf', 'cpg_float', '1e-4', True, None),
         ('eps_dual_inf', 'cpg_float', '1e-4', True, None), ('scaled_termination', 'cpg_int', '0', True, None),
         ('check_termination', 'cpg_int', '25', True, None), ('warm_starting', 'cpg_int', '1', True, 'warm_start'),
         ('verbose', 'cpg_int', '0', False, None), ('polishing', 'cpg_int', '0', False, None),
         ('polish_refine_iter', 'cpg_int', '0', False, None), ('delta', 'cpg_float', '1e-6', False, None)]


class ADMMCUDAInterface(*_BASES):
    solver_name = 'ADMM-CUDA'
    cvxpy_solver_name = 'OSQP'          # canonicalise through cvxpy's OSQP (QP) path
    solver_type = 'quadratic'
    supports_quad_obj = True
    canon_p_ids = ['P', 'q', 'd', 'A', 'l', 'u']
    canon_p_ids_constr_vec = ['l', 'u']
    dual_var_split = False
    dual_var_names = ['y']
    # same decision tree as OSQP's (cvxpygen/solvers/osqp.py:20-61): matrices -> re-equilibrate + refactor per instance on the
    # GPU (admm_matpar_kernel), vectors -> the kernels' prologue
    parameter_update_structure = {
        'PA': ParameterUpdateLogic(UpdatePendingLogic(['P', 'A'], '&&', ['P', 'A']), _mat_call(True, True)),
        'P': ParameterUpdateLogic(UpdatePendingLogic(['P']), _mat_call(True, False)),
        'A': ParameterUpdateLogic(UpdatePendingLogic(['A']), _mat_call(False, True)),
        'qlu': ParameterUpdateLogic(UpdatePendingLogic(['q', 'l', 'u'], '&&', ['ql', 'qu', 'lu']), _vec_call(1, 1, 1)),
        'ql': ParameterUpdateLogic(UpdatePendingLogic(['q', 'l'], '&&', ['q', 'l']), _vec_call(1, 1, 0)),
        'qu': ParameterUpdateLogic(UpdatePendingLogic(['q', 'u'], '&&', ['q', 'u']), _vec_call(1, 0, 1)),
        'lu': ParameterUpdateLogic(UpdatePendingLogic(['l', 'u'], '&&', ['l', 'u']), _vec_call(0, 1, 1)),
        'q': ParameterUpdateLogic(UpdatePendingLogic(['q']), _vec_call(1, 0, 0)),
        'l': ParameterUpdateLogic(UpdatePendingLogic(['l']), _vec_call(0, 1, 0)),
        'u': ParameterUpdateLogic(UpdatePendingLogic(['u']), _vec_call(0, 0, 1)),
    }
    solve_function_call = '{prefix}cpg_b200_shim_solve()'

    # header and source files
    header_files = ['"cpg_b200_shim.h"']
    cmake_headers = ['${CMAKE_CURRENT_SOURCE_DIR}/*.h', '${CMAKE_CURRENT_SOURCE_DIR}/*.cuh']
    cmake_sources = ['${CMAKE_CURRENT_SOURCE_DIR}/cpg_b200_shim.c']

    inmemory_preconditioning = False          # equilibration happens on the GPU / at generation time, never in Canon_Params

    # workspace: the shim's result block lives in cpg_workspace.{h,c} (declare_workspace / define_workspace), prefixed
    ws_statically_allocated_in_solver_code = False
    ws_ptrs = WorkspacePointerInfo(
        objective_value='cpg_b200_shim_info.obj_val',
        iterations='cpg_b200_shim_info.iter',
        status='cpg_b200_shim_info.status',
        primal_residual='cpg_b200_shim_info.prim_res',
        dual_residual='cpg_b200_shim_info.dual_res',
        primal_solution='sol_x',
        dual_solution='sol_{dual_var_name}')
    sol_statically_allocated = True           # CPG_Prim / CPG_Dual point into {prefix}sol_x / {prefix}sol_y
    status_is_int = False
    numeric_types = {'float': 'double', 'int': 'int'}

    # solver settings
    stgs_dynamically_allocated = False
    stgs_requires_extra_struct_type = False
    stgs_direct_write_ptr = '(&{prefix}cpg_b200_shim_settings)'
    stgs_reset_function = {'name': 'cpg_b200_shim_default_settings', 'ptr': '&{prefix}cpg_b200_shim_settings'}
    stgs = {n: Setting(t, d, en, cv) for n, t, d, en, cv in _STGS}
    docu = 'DESIGN.md'

    def __init__(self, data=None, p_prob=None, enable_settings=(), family=None):
        self.stgs = {n: Setting(t, d, en, cv) for n, t, d, en, cv in _STGS}     # per instance: configure_settings mutates it
        if family is None:                        # reference-style construction (needs cvxpygen)
            super().__init__(data, p_prob, list(enable_settings))
        else:                                     # cvxpy-free construction
            self.n_var, self.n_eq, self.n_ineq = family.n_var, family.n_eq, family.n_ineq
            self.enable_settings = list(enable_settings)
            self.canon_constants = {}
            for s in self.enable_settings:
                if s in self.stgs:
                    self.stgs[s].enabled = True
        self.family = family
        self.setup = None

    # ---- the settings views of SolverInterface (cvxpygen/solvers/_interface.py:183-197)
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
        pass

    @staticmethod
    def ret_prim_func_exists(variable_info) -> bool:      # _interface.py:120-122
        return any(variable_info.sym) or any(s == 1 for s in variable_info.name_to_size.values())

    @staticmethod
    def ret_dual_func_exists(dual_variable_info) -> bool:  # _interface.py:124-126
        return any(s == 1 for s in dual_variable_info.name_to_size.values())

    # ---- build-system hooks (cvxpygen/solvers/_interface.py:203-236; OSQP's: cvxpygen/solvers/osqp.py:148-170)
    def cmake_context_extra(self) -> dict:
        sdir = '${CMAKE_CURRENT_SOURCE_DIR}/solver_code'
        return {'solver_code_cmake_include_dir': sdir, 'extra_cmake_include_dirs': [sdir], 'packages': ['CUDAToolkit'],
                'cmake_target_link_libs': ['${CMAKE_CURRENT_SOURCE_DIR}/../libcpg_b200.so', 'CUDA::cudart'], 'cmake_definitions': []}

    def setup_py_context(self) -> dict:
        return {'solver_code_include_dir': "os.path.join('c', 'solver_code')", 'extra_solver_include_dirs': [],
                'extra_cpp_include_dirs': ["os.path.join('c', 'solver_code')"], 'extra_lib_names_windows': None,
                'extra_lib_names_unix': ['cpg_b200'], 'extra_objects': ["os.path.join('libcpg_b200.so')"], 'license': 'Apache 2.0'}

    # ---- workspace hooks: called by write_workspace_prot / write_workspace_def when the workspace is not the solver's own
    #      (cvxpygen/utils.py:661-662, 862-863)
    def declare_workspace(self, f, prefix, parameter_canon) -> None:
        m = self.n_eq + self.n_ineq
        f.write('\n// ADMM-CUDA workspace: canonical solution, solver info and settings of the last cpg_solve()\n')
        f.write(f'extern cpg_float {prefix}sol_x[{self.n_var}];\n')
        f.write(f'extern cpg_float {prefix}sol_y[{max(m, 1)}];\n')
        f.write(f'extern CpgB200ShimInfo {prefix}cpg_b200_shim_info;\n')
        f.write(f'extern CpgB200ShimSettings {prefix}cpg_b200_shim_settings;\n')

    def define_workspace(self, f, prefix, parameter_canon) -> None:
        m = self.n_eq + self.n_ineq
        f.write('\n// ADMM-CUDA workspace\n')
        f.write(f'cpg_float {prefix}sol_x[{self.n_var}];\n')
        f.write(f'cpg_float {prefix}sol_y[{max(m, 1)}];\n')
        f.write(f'CpgB200ShimInfo {prefix}cpg_b200_shim_info = {{0, 0, "unsolved", 0, 0, -10}};\n')
        f.write(f'CpgB200ShimSettings {prefix}cpg_b200_shim_settings = {{'
                + ', '.join(d for _, _, d, _, _ in _STGS) + '};\n')

    # ---- solver code generation (cvxpygen/generator.py:124-146 calls this before the writer runs)
    def generate_code(self, configuration, code_dir, solver_code_dir, cvxpygen_directory, canon, gradient, prefix,
                      compile=False):
        """Writes <code_dir>/c/solver_code: the CUDA sources of the canonical-level family + the shim.  `compile=True`
        also runs nvcc (the reference compiles later, in its own build step)."""
        from .. import codegen
        from ..ir import CanonFamily
        from ..offline.qp_setup import setup_qp_family
        from ..shim import write_shim
        pc = canon.parameter_canon
        changes = {k: bool(v) for k, v in pc.p_id_to_changes.items()}
        import numpy as np
        # a matrix is a per-instance input only if a user parameter enters it: the other one keeps OSQP's semantics of a NULL
        # argument to osqp_update_data_mat (cvxpygen/solvers/osqp.py:20-33)
        mats = tuple(k for k in ('P', 'A') if changes.get(k))
        name = getattr(self.family, 'name', None) or os.path.basename(os.path.abspath(code_dir))
        clip = lambda v: np.clip(np.asarray(v, dtype=float), -1e30, 1e30)     # the writer emits +-inf as +-1e30 too (utils.replace_inf)
        fam = CanonFamily.from_canonical_qp(name + '_canonical', pc.p['P'], pc.p['q'], pc.p['A'], clip(pc.p['l']), clip(pc.p['u']),
                                            n_eq=self.n_eq, matrix_params=mats)
        fam.is_maximization = False      # the sign flip is applied by the emitted cpg_retrieve_info (cvxpygen/utils.py:980)
        self.setup = setup_qp_family(fam, ['q', 'l', 'u'] + list(mats))
        cprefix = prefix or ''
        codegen.write_solver_sources(self.setup, solver_code_dir, prefix=cprefix)
        write_shim(self.setup, solver_code_dir, cprefix, matrices=mats)
        if compile:
            codegen.compile_solver_sources(solver_code_dir, os.path.join(code_dir, 'libcpg_b200.so'))
        return self.setup
