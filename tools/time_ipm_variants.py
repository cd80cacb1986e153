"""A/B of the gather-plan cell-size model of the IPM-CUDA kernel (offline/gather.py:_cost) on config 3's family.
   python tools/time_ipm_variants.py build      # here: one library per (c_fixed, c_shuf) under tools/_variants/ipm_*
   python tools/time_ipm_variants.py run        # on the GPU box"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VDIR = os.path.join(ROOT, 'tools', '_variants')
GRID = [(10, 3), (20, 3), (5, 3), (10, 6), (10, 1.5), (30, 6), (3, 1.5), (20, 1.5)]

if sys.argv[1] == 'build':
    procs = []
    for cf_, cs in GRID:
        d = os.path.join(VDIR, f'ipm_f{cf_}_s{cs}')
        code = ("from cvxpygen_b200 import families, cpg; cpg.generate_code(families.portfolio_socp(100, 10), code_dir=%r, solver='IPM-CUDA', "
                "batch_params=['a', 'w_prev'])" % d)
        procs.append(subprocess.Popen([sys.executable, '-c', code], env={**os.environ, 'PYTHONPATH': ROOT, 'CPG_GATHER_C_FIXED': str(cf_), 'CPG_GATHER_C_SHUF': str(cs)},
                                      stdout=subprocess.DEVNULL, stderr=subprocess.PIPE))
    for p in procs:
        _, err = p.communicate()
        if p.returncode:
            print(err.decode()[-500:])
    print(sorted(os.listdir(VDIR)))
else:
    import numpy as np, torch
    from cvxpygen_b200 import runtime
    import bench
    B = 20000
    P = torch.from_numpy(bench.WORKLOADS['portfolio_socp'].host_params(B, 1)).cuda()
    base = None
    for cf_, cs in GRID:
        d = os.path.join(VDIR, f'ipm_f{cf_}_s{cs}')
        if not os.path.exists(os.path.join(d, 'libcpg_b200.so')):
            continue
        mod = runtime.load(d).init()
        out = mod.solve_batch_device(P, return_canonical=True); torch.cuda.synchronize()
        ts = []
        for _ in range(2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); out = mod.solve_batch_device(P, out=out, return_canonical=True); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        it = out.iter.cpu().numpy(); x = out.sol_x.cpu().numpy()
        if base is None:
            base = (it, x)
        print(json.dumps(dict(c_fixed=cf_, c_shuf=cs, ms=float(np.median(ts)), inst_per_s=B / (np.median(ts) * 1e-3), mean_iter=float(it.mean()),
                              optimal=float((out.status.cpu().numpy() == 0).mean()), iter_equal_to_base=float((it == base[0]).mean()),
                              max_rel_x_vs_base=float((np.abs(x - base[1]).max(1) / np.abs(base[1]).max(1)).max()))), flush=True)
