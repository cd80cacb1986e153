#!/usr/bin/env python
"""A/B timing of compile-time kernel variants (extra nvcc -D flags) of a standard family.
   python tools/time_variants.py build                 # here (no GPU): generate + compile tools/_variants/<variant>
   python tools/time_variants.py run  [--grad]         # on the GPU box: device-resident timing of each variant
Variants are listed in VARIANTS: name -> (family, extra nvcc flags, batch)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
VDIR = os.path.join(ROOT, 'tools', '_variants')
VARIANTS = {
    'ltv_atomic': ('mpc_ltv_12_4_10', '', 20000),
    'ltv_gather': ('mpc_ltv_12_4_10', '-DCPG_TAIL_GATHER_FACTOR=1', 20000),
    'mpc_atomic': ('mpc_12_4_10', '', 100000),
    'mpc_gather': ('mpc_12_4_10', '-DCPG_TAIL_GATHER_FACTOR=1', 100000),
}


def build(names):
    import concurrent.futures as cf
    from cvxpygen_b200 import standard, cpg

    def one(v):
        famname, flags, _ = VARIANTS[v]
        fam_fn, batch = standard.STANDARD[famname]
        d = os.path.join(VDIR, v)
        cpg.generate_code(fam_fn(), code_dir=d, batch_params=batch, wrapper=False)
        from cvxpygen_b200 import codegen
        codegen.compile_code(d, extra_flags=flags.split())
        return d
    os.makedirs(VDIR, exist_ok=True)
    with cf.ThreadPoolExecutor(4) as ex:
        for d in ex.map(one, names):
            print('built', d)


def run(names, reps=3):
    import numpy as np, torch
    from cvxpygen_b200 import runtime, standard
    from helpers import ltv_batch
    for v in names:
        famname, flags, B = VARIANTS[v]
        d = os.path.join(VDIR, v)
        if not os.path.exists(os.path.join(d, 'libcpg_b200.so')):
            continue
        mod = runtime.Module(d).init()
        fam = standard.STANDARD[famname][0]()
        if mod.has_matrix_params:
            P = torch.from_numpy(mod.pack_params(ltv_batch(fam, B, seed=31))).cuda()
        else:
            P = torch.from_numpy(np.random.default_rng(1).uniform(-1, 1, (B, mod.dims.n_param))).cuda()
        out = mod.solve_batch_device(P, return_canonical=True)
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); mod.solve_batch_device(P, out=out); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        rec = dict(variant=v, flags=flags, B=B, fwd_ms=float(np.median(ts)), fwd_inst_per_s=B / (np.median(ts) * 1e-3),
                   mean_iter=float(out.iter.float().mean()), frac_solved=float((out.status == 1).float().mean()))
        dprim = torch.randn((B, mod.dims.n_prim), dtype=torch.float64, device='cuda')
        g = (lambda dp=None: mod.gradient_batch_device_mat(P, out.sol_x, out.sol_y, dprim, dparams=dp)) if mod.has_matrix_params \
            else (lambda dp=None: mod.gradient_batch_device(out.sol_y, dprim, dparams=dp))
        dp = g(); torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g(dp); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        rec.update(bwd_ms=float(np.median(ts)), bwd_inst_per_s=B / (np.median(ts) * 1e-3))
        print(json.dumps(rec), flush=True)


if __name__ == '__main__':
    names = [a for a in sys.argv[2:] if a in VARIANTS] or list(VARIANTS)
    build(names) if sys.argv[1] == 'build' else run(names)
