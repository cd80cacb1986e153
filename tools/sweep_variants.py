#!/usr/bin/env python
"""Build-and-time sweep of compile-time variants of the MPC solver library (warps per CTA, ...).
   python tools/sweep_variants.py build     # here (no GPU): generates + compiles tools/_variants/*
   python tools/sweep_variants.py run       # on the GPU box: times each variant on a 100k batch"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VDIR = os.path.join(ROOT, 'tools', '_variants')
VARIANTS = {'ni2_w12': dict(ni=2, warps=12), 'ov3': dict(ni=2, warps=12), 'ov5': dict(ni=2, warps=12), 'ov12': dict(ni=2, warps=12), 'ov16': dict(ni=2, warps=12)}

def build():
    from cvxpygen_b200 import families, cpg
    os.makedirs(VDIR, exist_ok=True)
    import concurrent.futures as cf
    with cf.ThreadPoolExecutor(6) as ex:
        list(ex.map(lambda kv: os.path.exists(os.path.join(VDIR, kv[0], 'libcpg_b200.so')) or cpg.generate_code(families.mpc(12, 4, 10), code_dir=os.path.join(VDIR, kv[0]), batch_params=['x_init'],
                                                 solver_opts=kv[1], verbose=True), VARIANTS.items()))

def run(B=100000, reps=5):
    import numpy as np, torch
    from cvxpygen_b200 import runtime
    xi = torch.from_numpy(np.random.default_rng(1).uniform(-1, 1, (B, 12))).cuda()
    for name in sorted(os.listdir(VDIR)):          # every built variant directory (also those built by hand for a sweep)
        d = os.path.join(VDIR, name)
        if not os.path.exists(os.path.join(d, 'libcpg_b200.so')):
            continue
        try:
            mod = runtime.Module(d).init()
            out = None
            for _ in range(2):
                out = mod.solve_batch_device(xi, out=out)
        except Exception as e:        # e.g. a warp count whose register allocation does not fit (warps are allocated four at a time)
            print(json.dumps(dict(variant=name, error=str(e)[:200])))
            continue
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
