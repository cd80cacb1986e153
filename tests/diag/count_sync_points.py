#!/usr/bin/env python
"""Dependent-chain census of the warp-per-instance kernels on the SIMT emulator: shuffles and warp barriers executed per
instance and per ADMM iteration (per warp = per-lane count, every lane executes the same ones).  Each of them is one step
of the warp's serial chain -- the quantity that bounds a latency-bound kernel.  Usage (repo root, CPU only):
    python tests/diag/count_sync_points.py [mpc_ltv_12_4_10] [n_instances]"""
import ctypes as C
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from cvxpygen_b200 import families, standard          # noqa: E402
from test_simt_emulation import build_emu, run_solve, _rows   # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else 'mpc_ltv_12_4_10'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
fam_fn, batch = standard.STANDARD[name]
fam = fam_fn()
st, lib, dims = build_emu(fam, batch, tempfile.mkdtemp())
cnt = (C.c_longlong * 3)()
matpar = bool(dims[5])
if matpar:
    params = families.mpc_ltv_batch(fam, B, seed=31) if name.startswith('mpc_ltv') else families.mpc_reference_batch(fam, B, seed=1)
    rows = _rows(fam, st, params, B)
    fn = 'emu_matpar_solve'
else:
    rows = np.random.default_rng(1).uniform(-1, 1, (B, dims[2]))
    fn = 'emu_tail_solve'
lib.emu_counters(cnt)
out = run_solve(lib, fn, dims, rows, grid=1)
lib.emu_counters(cnt)
warps_lanes = 32
ex, sw = cnt[0] / warps_lanes, cnt[1] / warps_lanes          # per warp (idle warps of the block execute none)
iters = int(out['iter'].sum())
print(json.dumps(dict(family=name, kernel=fn, instances=B, iterations=iters, shuffles_per_instance=ex / B, syncwarps_per_instance=sw / B,
                      shuffles_per_iteration=ex / iters, syncwarps_per_iteration=sw / iters,
                      nnz_L=st.stats['nnz_L'], levels=st.stats['n_levels'], fwd_tiles=len(st.refactor.fwd_tiles), bwd_tiles=len(st.refactor.bwd_tiles))))
