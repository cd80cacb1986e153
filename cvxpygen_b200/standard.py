"""The problem families this repository builds ahead of time (BASELINE.json `configs`).

`build_all()` is what `__graft_entry__.build()` runs: generate + nvcc-compile every family into
cvxpygen_b200/_generated/<name>/ (in-tree, so the .so files travel to the GPU box).
"""
import os
from typing import Dict, Tuple, Callable, List

from . import families
from .cpg import generate_code

GENERATED_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_generated')

# name -> (family builder, batched parameters)
STANDARD: Dict[str, Tuple[Callable, List[str]]] = {
    'mpc_12_4_10': (lambda: families.mpc(12, 4, 10), ['x_init']),          # BASELINE config 2 / 5 (headline)
    'mpc_6_3_10': (lambda: families.mpc(6, 3, 10), ['x_init']),             # the reference test's MPC size
    'mpc_6_3_10_dmma': (lambda: families.mpc(6, 3, 10), ['x_init']),        # the same through the opt-in FP64 tensor-core main kernel
    'nonneg_LS_3_2': (lambda: families.nonneg_ls(3, 2), ['b']),             # BASELINE config 1 (README example)
    'random_qp_20_5_15': (lambda: families.random_qp(20, 5, 15), ['q', 'b', 'h']),  # unstructured sparsity, q/l/u all batched
    'portfolio_qp_50_10': (lambda: families.portfolio_qp(50, 10), ['a', 'w_prev']),   # the reference's portfolio test problem (QP form, OSQP)
    'box_qp_6_8': (lambda: families.box_qp(6, 8), ['q', 'l', 'u']),          # corner cases: type changes, infeasibility
    # f3: a family whose tile schedule (530 KB) is larger than one SM's shared memory -- 1 500-row KKT, nnz(L) = 6 777; solved by the
    # per-instance-factor kernel alone with its tables read through L2 (codegen.py: CPG_FAM_BIG)
    'random_qp_700_100_700': (lambda: families.random_qp(700, 100, 700, density=0.002, seed=1), ['q', 'b', 'h']),
    # f2: the MPC family with its matrices as per-instance parameters (dynamics A, B and diagonal stage costs)
    'mpc_ltv_6_3_10': (lambda: families.mpc_ltv(6, 3, 10), ['A', 'B', 'qdiag', 'rdiag', 'x_init']),
    'mpc_ltv_12_4_10': (lambda: families.mpc_ltv(12, 4, 10), ['A', 'B', 'qdiag', 'rdiag', 'x_init']),
    'mpc_ref_6_3_10': (lambda: families.mpc_reference(10), ['Psqrt', 'Qsqrt', 'Rsqrt', 'A', 'B', 'x_init']),   # the reference's test MPC, all parameters
    'actuator_1_3': (lambda: families.actuator(), ['A', 'w', 'lamb_sm', 'kappa', 'u_prev', 'u_min', 'u_max']),   # degenerate shapes, scalar parameter in P
    'osqp_update_matrices_5_8': (lambda: families.osqp_update_matrices_kat()[0], ['q', 'l', 'u', 'P', 'A']),   # OSQP's own KAT for matrix updates
    'nonneg_LS_3_2_A': (lambda: families.nonneg_ls(3, 2, name='nonneg_LS_3_2_A'), ['A', 'b']),   # README example, A per instance
    'portfolio_socp_100_10': (lambda: families.portfolio_socp(100, 10), ['a', 'w_prev']),   # BASELINE config 3 (IPM-CUDA)
    # per-instance MATRIX parameters on the conic path: the factor loadings F (in A) and d_sqrt (in G) of the portfolio problem are
    # user parameters in the reference's example (examples/portfolio.ipynb); the kernel re-equilibrates per instance
    'portfolio_socp_mat_100_10': (lambda: families.portfolio_socp(100, 10, matrix_params=True), ['a', 'w_prev', 'F', 'd_sqrt']),
    'portfolio_socp_mat_20_4': (lambda: families.portfolio_socp(20, 4, matrix_params=True), ['a', 'w_prev', 'F', 'd_sqrt']),
    # generic conic families (every vector batched): three cones + equalities, and a pure LP; exit flags 0 / 1 / 2
    'random_socp_30_8_20_3x5x4': (lambda: families.random_socp(30, 8, 20, (3, 5, 4), seed=5), ['c', 'b', 'h']),
    'random_socp_20_5_30_lp': (lambda: families.random_socp(20, 5, 30, (), seed=6), ['c', 'b', 'h']),
    # the reference's network-flow LP (tests/test_E2E_LP.py, solved there with ECOS): vectors per instance, routing matrix shared
    'adp_socp_6_3': (lambda: families.adp_socp(), ['f']),      # the reference's SOCP test problem (tests/test_E2E_SOCP.py), four second-order cones, no LP cone
    'network_lp_50_10': (lambda: families.network_lp(50, 10), ['c', 'w', 'f_min', 'f_max']),
    # f4: the reference's two-stage gradient (QP family, conic forward solve by IPM-CUDA, QP backward pass): <dir> + <dir>/gradient
    'mpc_6_3_10_two_stage': (lambda: families.mpc(6, 3, 10), ['x_init']),
}
TWO_STAGE_NAMES: List[str] = ['mpc_6_3_10_two_stage']
SOLVER_OPTS: Dict[str, dict] = {'mpc_6_3_10_dmma': {'dmma': True}}      # non-default code-generation options of a standard family

# families solved by the ADMM (QP) backend / by the interior-point (SOCP) backend
SOCP_NAMES: List[str] = [n for n in STANDARD if '_socp_' in n or n.startswith('network_lp') or n.endswith('_two_stage')]     # conic families (IPM-CUDA)
MATPAR_NAMES: List[str] = ['mpc_ltv_6_3_10', 'mpc_ltv_12_4_10', 'mpc_ref_6_3_10', 'actuator_1_3', 'osqp_update_matrices_5_8',
                           'nonneg_LS_3_2_A']
BIG_NAMES: List[str] = ['random_qp_700_100_700']
QP_NAMES: List[str] = [n for n in STANDARD if n not in SOCP_NAMES and n not in MATPAR_NAMES and n not in BIG_NAMES and n not in SOLVER_OPTS]


def code_dir(name: str) -> str:
    return os.path.join(GENERATED_DIR, name)


def build(name: str, force: bool = False, verbose: bool = False) -> str:
    d = code_dir(name)
    if not force and os.path.exists(os.path.join(d, 'libcpg_b200.so')):
        return d
    fam_fn, batch = STANDARD[name]
    os.makedirs(GENERATED_DIR, exist_ok=True)
    fam = fam_fn()
    two_stage = name in TWO_STAGE_NAMES
    solver = 'IPM-CUDA' if (fam.solver_type == 'conic' or two_stage) else 'ADMM-CUDA'
    generate_code(fam, code_dir=d, solver=solver, batch_params=batch, prefix='', wrapper=True, verbose=verbose, gradient=two_stage,
                  solver_opts=SOLVER_OPTS.get(name))
    return d


def build_all(force: bool = False, verbose: bool = False, jobs: int = None):
    """Generate + compile every standard family.  The nvcc invocations (minutes of front-end time for the families with
    long generated straight-line solves) run concurrently, one subprocess per family, on the host cores."""
    import concurrent.futures as cf
    jobs = jobs or max(1, min(len(STANDARD), (os.cpu_count() or 2), 8))      # nvcc's front end takes ~1-2 GB per family
    with cf.ThreadPoolExecutor(jobs) as ex:
        order = sorted(STANDARD, key=lambda n: (n not in BIG_NAMES, n not in ('mpc_ltv_12_4_10', 'mpc_12_4_10', 'portfolio_qp_50_10')))   # longest compiles first
        futs = {name: ex.submit(build, name, force, verbose) for name in order}
        return {name: futs[name].result() for name in STANDARD}


def load(name: str, device: int = 0):
    from . import runtime
    if name in TWO_STAGE_NAMES:
        from .two_stage import TwoStageModule
        return TwoStageModule(build(name), device)
    return runtime.load(build(name), device)
