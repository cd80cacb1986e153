"""`ADMMCUDAInterface` -- the solver-plugin class of the ADMM-CUDA backend (boundary b1).

It exposes the attribute set the reference's writer reads from a `SolverInterface`
(reference: cvxpygen/solvers/_interface.py:82-258; the OSQP instance of it is
cvxpygen/solvers/osqp.py:16-118) so that it can be registered next to `OSQPInterface`
(see INTEGRATION.md section 1).  When the reference package is importable the class derives from its
`QPCanonMixin` + `SolverInterface`; otherwise (this container: cvxpy is absent) it stands alone and is
constructed from a `CanonFamily`.  The class carries no solver arithmetic: `generate_code` runs the
offline setup and emits the CUDA sources.
"""
from dataclasses import dataclass

try:                                              # reference present: be a real plugin
    from cvxpygen.solvers import SolverInterface as _RefBase, QPCanonMixin as _RefMixin   # pragma: no cover
    _BASES = (_RefMixin, _RefBase)
except Exception:                                 # cvxpy / cvxpygen absent
    _BASES = (object,)


@dataclass
class Setting:
    """Same fields as the reference's Setting (cvxpygen/mappings.py:139-145)."""
    type: str
    default: str
    enabled: bool = True
    name_cvxpy: str = None


class ADMMCUDAInterface(*_BASES):
    solver_name = 'ADMM-CUDA'
    cvxpy_solver_name = 'OSQP'          # canonicalise through cvxpy's OSQP (QP) path
    solver_type = 'quadratic'
    supports_quad_obj = True
    canon_p_ids = ['P', 'q', 'd', 'A', 'l', 'u']
    canon_p_ids_constr_vec = ['l', 'u']
    dual_var_split = False
    dual_var_names = ['y']
    # vectors are canonicalised inside the kernel per instance; matrices at (re-)setup time on the host
    parameter_update_structure = {}
    solve_function_call = '{prefix}cpg_solve_batch_host(1, ...)'
    header_files = ['"cpg_b200.h"']
    cmake_headers, cmake_sources = [], []
    inmemory_preconditioning = False
    ws_statically_allocated_in_solver_code = True
    sol_statically_allocated = False
    status_is_int = False
    numeric_types = {'float': 'double', 'int': 'int'}
    stgs_dynamically_allocated = False
    stgs_requires_extra_struct_type = False
    stgs_direct_write_ptr = None
    stgs_reset_function = {'name': 'cpg_b200_default_settings', 'ptr': None}
    stgs = {                              # mirrors cvxpygen/solvers/osqp.py:102-115
        'max_iter': Setting('cpg_int', '4000'),
        'eps_abs': Setting('cpg_float', '1e-3'),
        'eps_rel': Setting('cpg_float', '1e-3'),
        'eps_prim_inf': Setting('cpg_float', '1e-4'),
        'eps_dual_inf': Setting('cpg_float', '1e-4'),
        'scaled_termination': Setting('cpg_int', '0'),
        'check_termination': Setting('cpg_int', '25'),
        'warm_starting': Setting('cpg_int', '1', name_cvxpy='warm_start'),
        'verbose': Setting('cpg_int', '0', enabled=False),
        'polishing': Setting('cpg_int', '0', enabled=False),
        'polish_refine_iter': Setting('cpg_int', '0', enabled=False),
        'delta': Setting('cpg_float', '1e-6', enabled=False),
    }
    docu = 'DESIGN.md'

    def __init__(self, data=None, p_prob=None, enable_settings=(), family=None):
        if family is None:                        # reference-style construction (needs cvxpygen)
            super().__init__(data, p_prob, list(enable_settings))
        else:                                     # cvxpy-free construction
            self.n_var, self.n_eq, self.n_ineq = family.n_var, family.n_eq, family.n_ineq
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
        pass

    def generate_code(self, configuration, code_dir, solver_code_dir, cvxpygen_directory, canon, gradient, prefix,
                      batch_params=None, compile=True):
        from .. import codegen
        from ..ir import CanonFamily
        from ..offline.qp_setup import setup_qp_family
        # gradient=True needs nothing extra: every generated library carries the batched backward pass (cpg_gradient_batch_*)
        fam = self.family if self.family is not None else CanonFamily.from_reference_canon(
            getattr(configuration, 'code_dir', 'problem'), canon, self)
        setup = setup_qp_family(fam, batch_params)
        codegen.write_code(setup, code_dir, prefix=(prefix or '').rstrip('_'))
        if compile:
            codegen.compile_code(code_dir)
        return setup
