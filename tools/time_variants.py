#!/usr/bin/env python
"""A/B timing of compile-time kernel variants (extra nvcc -D flags) of a standard family.
   python tools/time_variants.py build                 # here (no GPU): generate + compile tools/_variants/<variant>
   python tools/time_variants.py run  [--grad]         # on the GPU box: device-resident timing of each variant
Variants are listed in VARIANTS: name -> (family, extra nvcc flags, batch)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
VDIR = os.path.join(ROOT, 'tools', '_variants')
VARIANTS = {      # name -> (family, extra nvcc flags, batch[, solver_opts])
    'big_pre8': ('random_qp_700_100_700', '-DCPG_TAIL_PRE=8', 4000),
    'big_pre16': ('random_qp_700_100_700', '-DCPG_TAIL_PRE=16', 4000),
    'ltv_cur': ('mpc_ltv_12_4_10', '', 20000),
    'ltv_pre4': ('mpc_ltv_12_4_10', '-DCPG_TAIL_PRE=4', 20000),
    'ltv_pre12': ('mpc_ltv_12_4_10', '-DCPG_TAIL_PRE=12', 20000),
    'ltv_pre16': ('mpc_ltv_12_4_10', '-DCPG_TAIL_PRE=16', 20000),
    'ltv_pre20': ('mpc_ltv_12_4_10', '-DCPG_TAIL_PRE=20', 20000),
    'mpc_pre12': ('mpc_12_4_10', '-DCPG_TAIL_PRE=12', 100000, {'dmma': False}),
    'mpc_pre16': ('mpc_12_4_10', '-DCPG_TAIL_PRE=16', 100000, {'dmma': False}),
    'mpc_pre4': ('mpc_12_4_10', '-DCPG_TAIL_PRE=4', 100000, {'dmma': False}),
    'mpc_cur': ('mpc_12_4_10', '', 100000, {'dmma': False}),
    'ltv_form0': ('mpc_ltv_12_4_10', '-DCPG_TAIL_FACTOR_FORM=0', 20000),
    'ltv_nounroll': ('mpc_ltv_12_4_10', '-DCPG_EQ_UNROLL=0', 20000),
    'ltv_form0_nounroll': ('mpc_ltv_12_4_10', '-DCPG_TAIL_FACTOR_FORM=0 -DCPG_EQ_UNROLL=0', 20000),
    'ltv_atomic': ('mpc_ltv_12_4_10', '', 20000),
    'ltv_gather': ('mpc_ltv_12_4_10', '-DCPG_TAIL_GATHER_FACTOR=1', 20000),
    'mpc_atomic': ('mpc_12_4_10', '', 100000, {'dmma': False}),
    'mpc_gather': ('mpc_12_4_10', '-DCPG_TAIL_GATHER_FACTOR=1', 100000, {'dmma': False}),
    'mpc_dmma_g3': ('mpc_12_4_10', '', 100000, {'dmma': True, 'dmma_groups': 3}),
    'mpc_dmma_g2': ('mpc_12_4_10', '', 100000, {'dmma': True, 'dmma_groups': 2}),
}


def build(names):
    import concurrent.futures as cf
    from cvxpygen_b200 import standard, cpg

    def one(v):
        famname, flags, _ = VARIANTS[v][:3]
        opts = VARIANTS[v][3] if len(VARIANTS[v]) > 3 else None
        fam_fn, batch = standard.STANDARD[famname]
        d = os.path.join(VDIR, v)
        cpg.generate_code(fam_fn(), code_dir=d, batch_params=batch, wrapper=False, solver_opts=opts)
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
        famname, flags, B = VARIANTS[v][:3]
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
        kt = mod.kernel_times()
        rec = dict(variant=v, flags=flags, B=B, fwd_ms=float(np.median(ts)), fwd_inst_per_s=B / (np.median(ts) * 1e-3),
                   main_ms=kt['main'], tail_ms=kt['tail'], warps=int(mod.dims.warps_per_cta), smem=int(mod.dims.smem_bytes),
                   mean_iter=float(out.iter.float().mean()), frac_solved=float((out.status == 1).float().mean()))
        if not mod.has_matrix_params and famname == 'mpc_12_4_10' and '--parity' in sys.argv:
            import bench
            wl = bench.WORKLOADS['mpc']
            ph = P[:20000].cpu().numpy()
            r = mod.solve_batch(ph, return_canonical=True)
            ora = wl.reference(ph, os.cpu_count())
            rec['parity'] = dict(iter_equal=float((r.cpg_info.iter == ora['iter']).mean()), status_equal=float((r.cpg_info.status == ora['status']).mean()),
                                 max_rel_x=bench.relmax_rows(r.sol_x, ora['x']), max_rel_y=bench.relmax_rows(r.sol_y, ora['y']))
        dprim = torch.randn((B, mod.dims.n_prim), dtype=torch.float64, device='cuda')
        g = (lambda dp=None: mod.gradient_batch_device_mat(P, out.sol_x, out.sol_y, dprim, dparams=dp)) if mod.has_matrix_params \
            else (lambda dp=None: mod.gradient_batch_device(out.sol_y, dprim, dparams=dp))
        try:
            dp = g(); torch.cuda.synchronize()
            ts = []
            for _ in range(reps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); g(dp); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            rec.update(bwd_ms=float(np.median(ts)), bwd_inst_per_s=B / (np.median(ts) * 1e-3))
        except Exception as e:          # e.g. a family without a generated backward kernel
            rec.update(bwd_ms=None, bwd_note=str(e)[:80])
        print(json.dumps(rec), flush=True)


if __name__ == '__main__':
    names = [a for a in sys.argv[2:] if a in VARIANTS] or list(VARIANTS)
    build(names) if sys.argv[1] == 'build' else run(names)
