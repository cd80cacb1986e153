#!/usr/bin/env python
"""BASELINE config 5: MPC QP batch sweep on one GPU (device-resident inputs, CUDA events, 3 repetitions)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from cvxpygen_b200 import standard
mod = standard.load('mpc_12_4_10').init()
for B in (1000, 10000, 100000, 1000000):
    xi = torch.from_numpy(np.random.default_rng(1).uniform(-1, 1, (B, 12))).cuda()
    out = None
    for _ in range(2):
        out = mod.solve_batch_device(xi, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5 if B <= 100000 else 2
    e0.record()
    for _ in range(reps):
        out = mod.solve_batch_device(xi, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(json.dumps(dict(batch=B, ms=round(ms, 3), inst_per_s=round(B / ms * 1e3), solved=float((out.status == 1).float().mean()))))
