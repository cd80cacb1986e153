"""Code emission for the ADMM-CUDA backend: turns a `QPSetup` into a generated code directory.

Role of the reference's writer + compiler for this path (cvxpygen/writer.py:62-73,
cvxpygen/compiler.py:24-31): same directory skeleton (Appendix A of SURVEY.md), but the
"solver_code" is CUDA and the build step is one nvcc invocation for sm_100a.

    <code_dir>/
      __init__.py
      cpg_solver.py                   generated Python front end (cpg_solve, cpg_solve_batch)
      c/include/cpg_b200.h            C ABI (copy of include/cpg_b200.h)
      c/include/cpg_family.h          compile-time sizes of this family
      c/include/cpg_blob_layout.h     blob header struct (from offline/blob.py)
      c/include/cpg_workspace.h, cpg_solve.h    reference-compatible single-instance interface
      c/src/cpg_blob.c                constants blob as a static array (role of cpg_workspace.c)
      c/src/cpg_solve.c               cpg_update_<p> / cpg_solve / cpg_set_solver_* on a batch of one
      c/solver_code/admm_kernel.cuh, cpg_b200_module.cu        hand-written sm_100a sources
      libcpg_b200.so                  built in-tree
"""
import os
import shutil
import subprocess
import sys
import time
from typing import Optional

import numpy as np

from .offline.blob import header_struct_c, tail_header_struct_c, grad_header_struct_c, mat_header_struct_c, TAIL_HEADER_FIELDS
from .offline.qp_setup import QPSetup

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, 'csrc')
_ROOT_INCLUDE = os.path.join(os.path.dirname(_HERE), 'include')

SMEM_BUDGET = 232448 - 1024          # 227 KB opt-in limit per CTA minus static/reserved
# instances per warp of the main kernel and the matching cap on warps per CTA (registers: 65536 / (32 * warps))
DEFAULT_NI = 2
MAX_WARPS = {2: 12, 4: 8}   # NI=2: 12 x 168 regs measured faster than 14 x 128 (spills) and 10 x 200 (tools/sweep_variants.py)

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-Xcompiler', '-fPIC', '-shared']


def c_ident(prefix: str) -> str:
    """Prefix sanitising as in the reference (cvxpygen/generator.py:175-182)."""
    if prefix and not prefix[0].isalpha():
        prefix = f'_{prefix}'
    return f'{prefix}_' if prefix else ''


DMMA_MAX_GROUPS = 3         # 12 warps x <= 170 registers


def _family_header(setup: QPSetup, prefix: str, warps: Optional[int], ni: Optional[int] = None, dmma: Optional[bool] = None,
                   dmma_groups: Optional[int] = None, force_big: bool = False) -> str:
    nk = setup.n + setup.m
    w_stride = nk + (nk % 2)
    blob_pad = (len(setup.blob) + 127) // 128 * 128
    nb_slots = setup.stats.get('nb_slots', 0)
    if ni is None:
        ni = DEFAULT_NI
    assert ni in (2, 4)
    pair_stride = ni * w_stride + ni * nb_slots + 2
    # BIG families: the tile schedule of the main kernel does not fit next to one warp's work vectors (the reference generates
    # code for any size).  Such a family is solved by the per-instance-factor kernel alone (admm_tail_kernel: every instance is
    # queued at iteration 0, the warp factors K itself from the family's tables) with the constants read through L1 / L2
    # instead of being staged, so that shared memory holds only the warps' work vector + factor.
    big = bool(force_big) or blob_pad + pair_stride * 8 > SMEM_BUDGET       # force_big: exercise the path on small families (tests)
    if warps is None:
        warps = max(1, min(MAX_WARPS[ni], (SMEM_BUDGET - blob_pad) // (pair_stride * 8)))
    trail = max(setup.schedule.n_trailing_tiles, 0)
    s_stride = setup.refactor.n_slots + 32 + (setup.refactor.n_slots % 2)      # + one dummy slot per lane (padding ops of the factorisation)
    cblob_pad = (len(setup.blob_compact) + 127) // 128 * 128
    tail_stage = int(not big and cblob_pad + (w_stride + s_stride) * 8 <= SMEM_BUDGET)
    tail_warps = max(1, min(8, (SMEM_BUDGET - (cblob_pad if tail_stage else 0)) // ((w_stride + s_stride) * 8)))
    if (w_stride + s_stride) * 8 > SMEM_BUDGET:
        raise ValueError(f'one instance of this family (work vector + numeric factor = {(w_stride + s_stride) * 8} bytes) does not fit '
                         f'the {SMEM_BUDGET} bytes of shared memory of one SM')
    gblob_pad = (len(setup.grad_blob) + 127) // 128 * 128
    n_pad, m_pad = (setup.n + 1) // 2 * 2, (setup.m + 1) // 2 * 2
    grad_stride = s_stride + 2 * w_stride + n_pad + m_pad
    if setup.mat_blob:      # matrix-parameter backward pass: + canonical x and the entries of P per warp
        grad_stride += n_pad + (setup.nnzP + 2) // 2 * 2
    grad_ok = int(gblob_pad + grad_stride * 8 <= SMEM_BUDGET)       # else: the backward kernel is not generated for this size
    grad_warps = max(1, min(8, (SMEM_BUDGET - gblob_pad) // (grad_stride * 8))) if grad_ok else 1
    # f2: per-instance matrix parameters -- per warp in shared memory: w | S | Pv (nothing staged, tables stay in L2; the
    # scaled entries of A and the scalings D, 1/D, E, 1/E sit in a per-warp slice of a global scratch buffer)
    matpar = 1 if setup.mat_blob else 0
    mat_a, mat_p = (setup.nnzA + 2) // 2 * 2, (setup.nnzP + 2) // 2 * 2
    mat_stride = w_stride + s_stride + mat_p
    mat_g_stride = mat_a + 2 * n_pad + 2 * m_pad
    mat_warps = max(1, min(8, SMEM_BUDGET // (mat_stride * 8)))
    if matpar and mat_stride * 8 > SMEM_BUDGET:
        raise ValueError('one instance of this family (factor + matrices) does not fit in shared memory')
    # FP64 tensor-core main kernel (admm_dmma_kernel): groups of four warps x eight instances; per group w8 (positions padded to a
    # multiple of 4, x 8 instances) + the staging buffer; chosen when its tables fit next to the compact blob with >= 2 groups
    dm_w8 = (nk + 3) // 4 * 4 * 8
    dm_stage = 4 * max(setup.dmma_rounds, 1) * 64
    dm_bv = 2 * nb_slots + 2
    dblob_pad = (len(setup.dmma_blob) + 127) // 128 * 128
    dm_fixed = cblob_pad + dblob_pad + 64
    dm_groups = min(dmma_groups or DMMA_MAX_GROUPS, (SMEM_BUDGET - dm_fixed) // ((dm_w8 + dm_stage + 4 * dm_bv) * 8)) if setup.dmma_blob else 0
    # opt-in (solver_opts={'dmma': True}): measured on B200 the tensor-core kernel is correct (identical iteration counts on 100 000
    # instances) but 28 % slower than the straight-line kernel on the MPC family -- its compressed coefficient tables cost as many
    # instructions and shared-memory wavefronts per useful FMA as the dense steps they replace
    use_dmma = int(bool(setup.dmma_blob) and dm_groups >= 2 and dmma is True and dm_w8 // 4 >= 2 * w_stride and not big)
    # tile headers of the per-instance triangular solves as a compile-time table (constant memory, admm_kernel.cuh:kTailTiles)
    th = dict(zip([n_ for _, n_ in TAIL_HEADER_FIELDS], np.frombuffer(setup.tail_blob, dtype='<i4', count=len(TAIL_HEADER_FIELDS))))
    n_tt = max(int(th['n_fwd_tiles'] + th['n_bwd_tiles']), 1)
    tail_tiles = np.frombuffer(setup.tail_blob, dtype='<i4', count=8 * n_tt, offset=int(th['off_i32']) + 4 * int(th['i_tiles']))
    nlv = int(th['n_levels']) + 1
    tail_levels = np.concatenate([np.frombuffer(setup.tail_blob, dtype='<i4', count=nlv, offset=int(th['off_i32']) + 4 * int(th[k]))
                                  for k in ('i_level_ptr', 'i_cround_ptr', 'i_scale_ptr')])
    lines = [
        '/* Auto-generated by cvxpygen_b200 %s -- compile-time sizes of problem family "%s". */' % (time.strftime('%Y-%m-%d'), setup.family.name),
        '#ifndef CPG_FAMILY_H', '#define CPG_FAMILY_H',
    ]
    if prefix:
        lines.append(f'#define CPG_B200_PREFIX {prefix}')
    lines += [
        f'#define CPG_FAM_N {setup.n}', f'#define CPG_FAM_M {setup.m}', f'#define CPG_FAM_NPB {setup.npb}',
        f'#define CPG_FAM_TRAIL_TILES {trail}', f'#define CPG_FAM_WARPS {warps}',
        f'#define CPG_FAM_BLOB_BYTES {len(setup.blob)}', f'#define CPG_FAM_BLOB_BYTES_PAD {blob_pad}',
        f'#define CPG_FAM_CBLOB_BYTES_PAD {cblob_pad}', f'#define CPG_FAM_GBLOB_BYTES_PAD {gblob_pad}',
        f'#define CPG_FAM_GRAD_WARPS {grad_warps}', f'#define CPG_FAM_GRAD_STRIDE {grad_stride}',
        f'#define CPG_FAM_W_STRIDE {w_stride}', f'#define CPG_FAM_SCALING {int(setup.scaling)}',
        f'#define CPG_FAM_NI {ni}', f'#define CPG_FAM_MULTI_STRIDE {pair_stride}', f'#define CPG_FAM_S_STRIDE {s_stride}', f'#define CPG_FAM_TAIL_WARPS {tail_warps}',
        f'#define CPG_FAM_MATPAR {matpar}', f'#define CPG_FAM_MAT_WARPS {mat_warps}', f'#define CPG_FAM_MAT_STRIDE {mat_stride}',
        f'#define CPG_FAM_MAT_A_STRIDE {mat_a}', f'#define CPG_FAM_MAT_P_STRIDE {mat_p}', f'#define CPG_FAM_MAT_G_STRIDE {mat_g_stride}',
        f"#define CPG_FAM_TAIL_WORD_SHIFT {int(th['pad0'])}",
        '#define CPG_FAM_TAIL_LEVELS {' + ', '.join(str(int(v)) for v in tail_levels) + '}',
        '#define CPG_FAM_TAIL_TILES {' + ', '.join(str(int(v)) for v in tail_tiles) + '}',
        f'#define CPG_FAM_BIG {int(big)}', f'#define CPG_FAM_TAIL_STAGE {tail_stage}', f'#define CPG_FAM_GRAD {grad_ok}',
        f'#define CPG_FAM_DMMA {use_dmma}', f'#define CPG_FAM_DM_GROUPS {max(dm_groups, 1)}', f'#define CPG_FAM_DBLOB_BYTES_PAD {dblob_pad}',
        f'#define CPG_FAM_DM_W8 {dm_w8}', f'#define CPG_FAM_DM_STAGE {dm_stage}', f'#define CPG_FAM_DM_BV {dm_bv}',
        '#endif', '']
    return '\n'.join(lines)


def _blob_c(setup: QPSetup) -> str:
    out = ['/* Auto-generated by cvxpygen_b200: constants blobs (role of the reference\'s cpg_workspace.c / OSQP workspace.c). */',
           '#include "cpg_family.h"', '#include "cpg_b200.h"']
    for sym, data in (('cpg_blob', setup.blob), ('cpg_tail_blob', setup.tail_blob), ('cpg_cblob', setup.blob_compact),
                      ('cpg_gblob', setup.grad_blob), ('cpg_gS0', np.asarray(setup.grad_S0, dtype='<f8').tobytes()),
                      ('cpg_mblob', setup.mat_blob or b'\0' * 16), ('cpg_dblob', setup.dmma_blob or b'\0' * 32)):
        words = np.frombuffer(data + b'\0' * ((-len(data)) % 8), dtype='<u8')
        out.append(f'const unsigned long long CPG_B200_FN({sym}_words)[] __attribute__((aligned(128))) = {{')
        for i in range(0, len(words), 4):
            out.append('  ' + ', '.join('0x%016xULL' % int(v) for v in words[i:i + 4]) + ',')
        out.append('};')
        out.append(f'const unsigned int CPG_B200_FN({sym}_nbytes) = {len(data)}u;')
    return '\n'.join(out) + '\n'


def write_code(setup: QPSetup, code_dir: str, prefix: str = '', warps: Optional[int] = None, ni: Optional[int] = None,
               dmma: Optional[bool] = None, dmma_groups: Optional[int] = None, force_big: bool = False) -> None:
    prefix = c_ident(prefix)
    shutil.rmtree(code_dir, ignore_errors=True)
    for sub in ('c/src', 'c/include', 'c/build', 'c/solver_code'):
        os.makedirs(os.path.join(code_dir, sub))
    inc, src, sol = (os.path.join(code_dir, 'c', d) for d in ('include', 'src', 'solver_code'))
    shutil.copyfile(os.path.join(_ROOT_INCLUDE, 'cpg_b200.h'), os.path.join(inc, 'cpg_b200.h'))
    for f in ('admm_kernel.cuh', 'admm_multi_kernel.cuh', 'grad_kernel.cuh', 'matpar_kernel.cuh', 'cpg_b200_module.cu'):
        shutil.copyfile(os.path.join(_CSRC, f), os.path.join(sol, f))
    with open(os.path.join(inc, 'cpg_family.h'), 'w') as f:
        f.write(_family_header(setup, prefix, warps, ni, dmma, dmma_groups, force_big))
    with open(os.path.join(inc, 'cpg_blob_layout.h'), 'w') as f:
        f.write('/* Auto-generated by cvxpygen_b200 from offline/blob.py:HEADER_FIELDS. */\n#pragma once\n' + header_struct_c() + tail_header_struct_c() + grad_header_struct_c() + mat_header_struct_c())
    with open(os.path.join(inc, 'cpg_kkt_solve_gen.cuh'), 'w') as f:
        f.write(setup.solve_source if '#define CPG_FAM_BIG 0' in open(os.path.join(inc, 'cpg_family.h')).read()
                else '// BIG family: the straight-line schedule is not compiled (solved by the per-instance-factor kernel)\n')
    with open(os.path.join(src, 'cpg_blob.c'), 'w') as f:
        f.write(_blob_c(setup))
    with open(os.path.join(code_dir, 'cpg_blob.bin'), 'wb') as f:     # same bytes, for cpg_b200_load_constants / NCCL broadcast
        f.write(setup.blob)
    with open(os.path.join(code_dir, 'cpg_tail_blob.bin'), 'wb') as f:
        f.write(setup.tail_blob)
    import pickle
    with open(os.path.join(code_dir, 'cpg_family.pkl'), 'wb') as f:      # what update_shared_params needs to re-run the setup
        pickle.dump(dict(family=setup.family, batch_params=setup.batch_params, rho=setup.rho, sigma=setup.sigma,
                         scaling=setup.scaling, solve_source=setup.solve_source,
                         max_group_rows=getattr(setup, 'max_group_rows', 32), allow_trailing=getattr(setup, 'allow_trailing', True)), f)
    with open(os.path.join(code_dir, '__init__.py'), 'w') as f:
        f.write('')
    from .frontend import write_frontend
    write_frontend(setup, code_dir, prefix)


def write_solver_sources(setup: QPSetup, solver_code_dir: str, prefix: str = '', warps: Optional[int] = None,
                         ni: Optional[int] = None) -> None:
    """Everything libcpg_b200.so is compiled from, FLAT in one directory: what `SolverInterface.generate_code` puts into
    <code_dir>/c/solver_code when the reference's own generator drives the plugin (cvxpygen/generator.py:124-146; the
    OSQP plugin writes OSQP's embedded sources there, cvxpygen/solvers/osqp.py:120-146).  `prefix` is already a C identifier
    prefix (with trailing underscore) or empty."""
    os.makedirs(solver_code_dir, exist_ok=True)
    shutil.copyfile(os.path.join(_ROOT_INCLUDE, 'cpg_b200.h'), os.path.join(solver_code_dir, 'cpg_b200.h'))
    for f in ('admm_kernel.cuh', 'admm_multi_kernel.cuh', 'grad_kernel.cuh', 'matpar_kernel.cuh', 'cpg_b200_module.cu'):
        shutil.copyfile(os.path.join(_CSRC, f), os.path.join(solver_code_dir, f))
    with open(os.path.join(solver_code_dir, 'cpg_family.h'), 'w') as f:
        f.write(_family_header(setup, prefix, warps, ni))
    with open(os.path.join(solver_code_dir, 'cpg_blob_layout.h'), 'w') as f:
        f.write('/* Auto-generated by cvxpygen_b200 from offline/blob.py:HEADER_FIELDS. */\n#pragma once\n' + header_struct_c() + tail_header_struct_c() + grad_header_struct_c() + mat_header_struct_c())
    with open(os.path.join(solver_code_dir, 'cpg_kkt_solve_gen.cuh'), 'w') as f:
        f.write(setup.solve_source)
    with open(os.path.join(solver_code_dir, 'cpg_blob.c'), 'w') as f:
        f.write(_blob_c(setup))


def compile_solver_sources(solver_code_dir: str, out: str, verbose: bool = False, extra_flags=()) -> str:
    """nvcc (sm_100a) on a flat solver_code directory written by write_solver_sources -> shared library `out`."""
    srcs = [os.path.join(solver_code_dir, 'cpg_b200_module.cu'), os.path.join(solver_code_dir, 'cpg_blob.c')]
    env_flags = os.environ.get('CPG_B200_EXTRA_NVCC_FLAGS', '').split()
    cmd = ['nvcc'] + NVCC_FLAGS + list(extra_flags) + env_flags + ['-I', solver_code_dir] + srcs + ['-o', out]
    if verbose:
        sys.stdout.write(' '.join(cmd) + '\n')
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError('nvcc failed:\n' + res.stdout + res.stderr)
    return out


def lib_path(code_dir: str) -> str:
    return os.path.join(code_dir, 'libcpg_b200.so')


def compile_code(code_dir: str, verbose: bool = False, extra_flags=()) -> str:
    """One nvcc invocation for sm_100a, output in-tree (role of `setup.py build_ext --inplace`,
    cvxpygen/compiler.py:24-31)."""
    inc, src, sol = (os.path.join(code_dir, 'c', d) for d in ('include', 'src', 'solver_code'))
    out = lib_path(code_dir)
    srcs = [os.path.join(sol, 'cpg_b200_module.cu'), os.path.join(src, 'cpg_blob.c')]
    compat = os.path.join(src, 'cpg_solve.c')
    if os.path.exists(compat):
        srcs.append(compat)
    # CPG_B200_EXTRA_NVCC_FLAGS: kernel variants for A/B measurements on a GPU box, e.g. "-DCPG_TAIL_GATHER_FACTOR=1"
    env_flags = os.environ.get('CPG_B200_EXTRA_NVCC_FLAGS', '').split()
    cmd = ['nvcc'] + NVCC_FLAGS + list(extra_flags) + env_flags + ['-I', inc, '-I', sol] + srcs + ['-o', out]
    if verbose:
        cmd.insert(1, '-Xptxas'); cmd.insert(2, '-v')
        sys.stdout.write(' '.join(cmd) + '\n')
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError('nvcc failed:\n' + res.stdout + res.stderr)
    if verbose:
        sys.stdout.write(res.stdout + res.stderr)
    return out
