"""`IPMCUDAInterface` -- the solver-plugin class of the IPM-CUDA backend (boundary b1, conic families).

The conic twin of `ADMMCUDAInterface`: the attribute set the reference's writer reads from a `SolverInterface`
(reference: cvxpygen/solvers/_interface.py:82-258), with the values of the plugin it stands beside --
`ECOSInterface` (cvxpygen/solvers/ecos.py:16-134): conic canonical form (c, d, A, b, G, h) + cone dimensions, integer exit
flags, duals split into y (equalities) and z (cones), ECOS's settings table.  It carries no solver arithmetic:
`generate_code` runs the offline setup (offline/socp_setup.py) and emits the CUDA sources (codegen_ipm.py).
"""
from .admm_cuda import Setting

try:                                              # reference present: be a real plugin
    from cvxpygen.solvers import SolverInterface as _RefBase   # pragma: no cover
    _BASES = (_RefBase,)
except Exception:                                 # cvxpy / cvxpygen absent
    _BASES = (object,)


class IPMCUDAInterface(*_BASES):
    solver_name = 'IPM-CUDA'
    cvxpy_solver_name = 'ECOS'          # canonicalise through cvxpy's ECOS (conic) path
    solver_type = 'conic'
    supports_quad_obj = False
    canon_p_ids = ['c', 'd', 'A', 'b', 'G', 'h']
    canon_p_ids_constr_vec = ['b', 'h']
    dual_var_split = True
    dual_var_names = ['y', 'z']
    # the vectors are canonicalised inside the kernel per instance; matrices at (re-)setup time on the host
    parameter_update_structure = {}
    solve_function_call = '{prefix}cpg_socp_solve_batch_host(1, ...)'
    header_files = ['"cpg_b200_socp.h"']
    cmake_headers, cmake_sources = [], []
    inmemory_preconditioning = False    # equilibration happens offline (constants) and on chip (per-instance vectors)
    ws_statically_allocated_in_solver_code = True
    sol_statically_allocated = False
    status_is_int = True                # ECOS exit flags: 0 optimal, 1 / 2 primal / dual infeasible, +10 inaccurate, -1 maxit, ...
    numeric_types = {'float': 'double', 'int': 'int'}
    stgs_dynamically_allocated = False
    stgs_requires_extra_struct_type = False
    stgs_direct_write_ptr = None
    stgs_reset_function = {'name': 'cpg_socp_default_settings', 'ptr': None}
    stgs = {                              # mirrors cvxpygen/solvers/ecos.py:59-67
        'feastol': Setting('cpg_float', '1e-8'),
        'abstol': Setting('cpg_float', '1e-8'),
        'reltol': Setting('cpg_float', '1e-8'),
        'feastol_inacc': Setting('cpg_float', '1e-4'),
        'abstol_inacc': Setting('cpg_float', '5e-5'),
        'reltol_inacc': Setting('cpg_float', '5e-5'),
        'maxit': Setting('cpg_int', '100', name_cvxpy='max_iters'),
    }
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

    @property
    def stgs_names_enabled(self):
        return [n for n, s in self.stgs.items() if s.enabled]

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
