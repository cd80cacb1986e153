"""Public entry point, same shape as the reference's (cvxpygen/cpg.py:17-30):

    generate_code(problem, code_dir='cpg_code', solver=None, solver_opts=None,
                  enable_settings=[], prefix='', gradient=False, wrapper=True)

with two additions for the batched backend: `batch_params` (the user parameters that vary per
instance; the others are shared by the batch and folded into the constants at generation time)
and `problem` may be either a cvxpy Problem (needs cvxpy + cvxpygen importable: the reference's own
Canonicalizer is reused unchanged) or a `cvxpygen_b200.ir.CanonFamily` built without cvxpy.
"""
import sys
from typing import List, Optional

from .ir import CanonFamily

SOLVERS = ('ADMM-CUDA', 'IPM-CUDA')


# the reference solver whose canonical form each backend consumes (QP form of OSQP / conic form of ECOS)
REFERENCE_FORM = {'ADMM-CUDA': 'OSQP', 'IPM-CUDA': 'ECOS'}


def _canonicalize_with_reference(problem, solver_opts, enable_settings, solver='ADMM-CUDA'):
    """cvxpy-present path: run the reference's canonicaliser for the form the backend consumes and convert to the IR."""
    try:
        from cvxpygen.canonicalizer import Canonicalizer      # reference package, if installed
    except ImportError as e:
        raise ImportError('generate_code was given a cvxpy Problem, which needs cvxpy and cvxpygen to be '
                          'importable for canonicalisation; pass a cvxpygen_b200.ir.CanonFamily instead') from e
    canon, interface = Canonicalizer(solver=REFERENCE_FORM[solver.upper()], solver_opts=solver_opts,
                                     enable_settings=enable_settings).canonicalize(problem)
    fam = CanonFamily.from_reference_canon('problem', canon, interface)
    if fam.solver_type == 'conic':           # cone sizes: ECOSInterface.canon_constants (cvxpygen/solvers/ecos.py:76-83)
        cc = interface.canon_constants
        fam.cone_dims = {'l': int(cc['l']), 'q': [int(v) for v in cc['q']]}
    return fam


def generate_code(problem, code_dir='cpg_code', solver=None, solver_opts=None, enable_settings=[],
                  prefix='', gradient=False, wrapper=True, batch_params: Optional[List[str]] = None,
                  verbose=False):
    from . import codegen
    from .offline.qp_setup import setup_qp_family
    solver = 'ADMM-CUDA' if solver is None else solver
    if solver.upper() not in SOLVERS:
        raise ValueError(f'Unsupported solver: {solver}.')       # same text as cvxpygen/canonicalizer.py:83
    # gradient=True needs nothing extra: every generated library carries the batched backward pass
    # (cpg_gradient_batch_*); the flag is accepted for signature compatibility with the reference.
    sys.stdout.write(f'Generating code with cvxpygen_b200 ({solver.upper()}, sm_100a) ...\n')
    fam = problem if isinstance(problem, CanonFamily) else _canonicalize_with_reference(problem, solver_opts, enable_settings, solver)
    if not fam.params:
        raise ValueError('Solution does not depend on parameters. Aborting code generation.')  # canonicalizer.py:98-99
    opts = dict(solver_opts or {})
    if solver.upper() == 'IPM-CUDA':
        # SOCP families: Mehrotra interior-point kernel (role of solver='ECOS' in the reference, cvxpygen/solvers/ecos.py)
        from . import codegen_ipm
        from .offline.socp_setup import setup_socp_family
        if fam.solver_type == 'quadratic':
            # a QP handed to the conic backend (the reference solves QPs with ECOS / SCS / Clarabel too).  With gradient=True this
            # is the reference's two-stage route (cvxpygen/canonicalizer.py:54-65): conic forward solve, QP backward pass.
            from . import two_stage
            if gradient:
                out = two_stage.generate_two_stage(fam, code_dir, batch_params, prefix=prefix, verbose=verbose, compile=wrapper)
                sys.stdout.write('cvxpygen_b200 finished generating code (two-stage gradient: IPM-CUDA forward, QP backward).\n')
                return out
            fam = two_stage.conic_family_of_qp(fam, batch_params)
        elif gradient:
            raise ValueError('gradient=True needs a problem that has a QP canonical form (extended DPP, cvxpygen/canonicalizer.py:338-345): '
                             'pass the QP family and solver=\'IPM-CUDA\' for the two-stage route')
        from .offline.socp_setup import DEFAULT_THREADS
        setup = setup_socp_family(fam, batch_params, threads=int(opts.get('threads') or DEFAULT_THREADS))
        codegen_ipm.write_ipm_code(setup, code_dir, prefix=prefix, threads=opts.get('threads'))
        sys.stdout.write('cvxpygen_b200 finished generating code.\n')
        if wrapper:
            sys.stdout.write('Compiling CUDA solver library (nvcc, sm_100a) ...\n')
            codegen_ipm.compile_ipm_code(code_dir, verbose=verbose)
            sys.stdout.write('cvxpygen_b200 finished compiling.\n')
        return setup
    if fam.solver_type != 'quadratic':
        raise ValueError('ADMM-CUDA handles the QP canonical form; use solver=\'IPM-CUDA\' for conic families')
    setup = setup_qp_family(fam, batch_params, rho=opts.get('rho', 0.1), sigma=opts.get('sigma', 1e-6),
                            scaling=opts.get('scaling', 10), max_group_rows=opts.get('max_group_rows', 32))
    codegen.write_code(setup, code_dir, prefix=prefix, warps=opts.get('warps'), ni=opts.get('ni'), dmma=opts.get('dmma'),
                       dmma_groups=opts.get('dmma_groups'), force_big=bool(opts.get('force_big')))
    sys.stdout.write('cvxpygen_b200 finished generating code.\n')
    if wrapper:
        sys.stdout.write('Compiling CUDA solver library (nvcc, sm_100a) ...\n')
        codegen.compile_code(code_dir, verbose=verbose)
        sys.stdout.write('cvxpygen_b200 finished compiling.\n')
    return setup
