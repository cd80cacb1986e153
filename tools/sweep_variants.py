#!/usr/bin/env python
"""Build-and-time sweep of compile-time variants of the MPC solver library (warps per CTA, ...).
   python tools/sweep_variants.py build     # here (no GPU): generates + compiles tools/_variants/*
   python tools/sweep_variants.py run       # on the GPU box: times each variant on a 100k batch"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VDIR = os.path.join(ROOT, 'tools', '_variants')
VARIANTS = {'ni2_w11': dict(ni=2, warps=11), 'ni2_w12': dict(ni=2, warps=12), 'ni2_w13': dict(ni=2, warps=13), 'ni2_w14': dict(ni=2, warps=14)}

def build():
    from cvxpygen_b200 import families, cpg
    os.makedirs(VDIR, exist_ok=True)
    for name, opts in VARIANTS.items():
        cpg.generate_code(families.mpc(12, 4, 10), code_dir=os.path.join(VDIR, name), batch_params=['x_init'], solver_opts=opts)

def run(B=100000, reps=5):
    import numpy as np, torch
    from cvxpygen_b200 import runtime
    xi = torch.from_numpy(np.random.default_rng(1).uniform(-1, 1, (B, 12))).cuda()
    for name in VARIANTS:
        d = os.path.join(VDIR, name)
        if not os.path.exists(os.path.join(d, 'libcpg_b200.so')):
            continue
        mod = runtime.Module(d).init()
        out = None
        for _ in range(2):
            out = mod.solve_batch_device(xi, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            out = mod.solve_batch_device(xi, out=out)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        st = out.status.cpu().numpy()
        print(json.dumps(dict(variant=name, ms=ms, inst_per_s=B / ms * 1e3, solved=float((st == 1).mean()), mean_iter=float(out.iter.float().mean()))))

if __name__ == '__main__':
    build() if sys.argv[1] == 'build' else run()
