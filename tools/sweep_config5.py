#!/usr/bin/env python
"""BASELINE config 5: MPC QP batch sweep 1k .. 1M at 1 / 2 / 4 / 8 GPUs of one node, STRONG scaling (the batch is fixed, the
product API shards it: cpg_solve_batch_host_multi -- one process, one host thread per device, contiguous shards, host buffers in,
host buffers out).  One JSON line per (devices, batch): wall time of the call (median of 3), instances/s, and for N = 1 the
device-resident kernel time for comparison.   python tools/sweep_config5.py > gpurun_out/r2_config5_sweep.jsonl"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from cvxpygen_b200 import standard
mod = standard.load('mpc_12_4_10').init()
ndev = torch.cuda.device_count()
PINNED = '--pinned' in sys.argv        # caller-owned pinned host buffers in and out (the devices store their result rows directly)
d = mod.dims
for B in (1000, 10000, 100000, 1000000):
    xi = np.random.default_rng(1).uniform(-1, 1, (B, 12))
    out = None
    if PINNED:
        pin = lambda shape, dt=torch.float64: torch.empty(shape, dtype=dt, pin_memory=True).numpy()
        xp = pin((B, 12)); xp[:] = xi; xi = xp
        out = dict(prim=pin((B, d.n_prim)), dual=pin((B, d.n_dual)), obj=pin(B), pri=pin(B), dua=pin(B), it=pin(B, torch.int32), st=pin(B, torch.int32))
    for N in (1, 2, 4, 8):
        if N > ndev:
            continue
        devs = list(range(N))
        mod.solve_batch_multi(xi, devices=devs, out=out)           # warm-up: contexts, staging buffers
        ts = []
        for _ in range(3):
            t0 = time.perf_counter(); r = mod.solve_batch_multi(xi, devices=devs, out=out); ts.append(time.perf_counter() - t0)
        t = float(np.median(ts))
        rec = dict(config='MPC QP (12,4,10) strong scaling through cpg_solve_batch_host_multi (%s host buffers)' % ('pinned' if PINNED else 'pageable'), n_gpus=N, batch=B,
                   ms=round(t * 1e3, 3), inst_per_s=round(B / t), frac_solved=float((r.cpg_info.status == 1).mean()))
        if N == 1:
            P = torch.from_numpy(np.ascontiguousarray(xi)).cuda(); dout = mod.solve_batch_device(P); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); mod.solve_batch_device(P, out=dout); e1.record(); torch.cuda.synchronize()
            rec['device_resident_ms'] = round(e0.elapsed_time(e1), 3)
        print(json.dumps(rec), flush=True)
