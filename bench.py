#!/usr/bin/env python
"""bench.py -- headline benchmark of the batched-solve hot path (BASELINE.json `metric`).

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a backend
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU path (oracle/_ref)

A *step* is one batched solve of `--batch` (default 100 000) MPC QP instances (n_x=12, n_u=4, N=10;
BASELINE.json configs[1]) per GPU with synthetic x_init ~ U[-1,1]^12 (seed 1 + rank).  Weak scaling:
every rank solves its own batch; value = all instances / max-over-ranks device time.

Timing: W >= 3 warm-up steps; every timed step is bracketed by CUDA events on the launching stream;
between steps a 256 MiB buffer is written to flush the 126 MB L2 (the flush is outside the event
pairs); the K event times are summed; max over ranks.  `e2e` is the same metric through the public
host-buffer API (pinned host memory: H2D of the parameters, kernel, D2H of every result array) timed
with the host clock around the synchronous call.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FAMILY = 'mpc_12_4_10'
WORKLOAD = 'MPC QP (n_x=12,n_u=4,N=10) batch=%d per GPU, ADMM-CUDA backend, OSQP default settings (eps 1e-3, adaptive rho)'
# --workload portfolio_socp: BASELINE.json configs[2] (portfolio SOCP n=100 assets, batch 50k, IPM-CUDA backend); the default
# run (no flag) is the headline MPC workload above.
SOCP_FAMILY = 'portfolio_socp_100_10'
SOCP_WORKLOAD = 'portfolio SOCP (n=100 assets, 10 factors; 512 vars, 111 eq, 715 cone rows) batch=%d per GPU, IPM-CUDA backend, ECOS default settings (tol 1e-8)'
SOCP_BYTES_PER_INSTANCE = 200 * 8 + (210 + 112) * 8 + 40        # a, w_prev in; w, delta_w, f + duals out; info (~4.2 KB)
SOCP_TRAFFIC_BYTES_PER_INSTANCE = 2164                           # ncu at batch 1184: (2.365 MB + 0.197 MB) / 1184
# algorithmic I/O and work per instance (SURVEY.md section 8d / DESIGN.md): 96 B in + 2.75 KB out + 40 B info
BYTES_PER_INSTANCE = 12 * 8 + (172 + 172) * 8 + 40
FLOP_PER_INSTANCE = 0.5e6
# dram__bytes_read.sum + dram__bytes_write.sum of admm_multi_kernel from the `ncu --set full` capture of this same command
# at batch 100000 (profiles/r1_v7_ncu_summary.md: 12.8 MB + 418.1 MB): 4.31 KB per instance.  The excess over the
# algorithmic 2.89 KB is register-spill write-back, not re-reads of inputs.
TRAFFIC_BYTES_PER_INSTANCE = 4309


def load_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.rows = index, threading.Event(), []
        self.active = threading.Event()           # samples are taken only while a timed region is running
        self.active.set()

    def run(self):
        q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
            'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'
        while not self.stop_flag.is_set():
            if not self.active.is_set():
                self.stop_flag.wait(0.02)
                continue
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), f'--query-gpu={q}', '--format=csv,noheader,nounits'],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(',')])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        if not self.rows:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['unsampled']}
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith('active') for r in self.rows)]
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': int(self.rows[0][1]) if self.rows[0][1].isdigit() else None,
                'reasons': reasons, 'samples': len(self.rows)}


def mpc_canonical_batch(B, seed):
    from cvxpygen_b200 import families
    fam = families.mpc(12, 4, 10)
    xi = np.random.default_rng(seed).uniform(-1, 1, (B, 12))
    l0, u0 = fam.canon_data('l'), fam.canon_data('u')
    L = np.tile(l0, (B, 1)); U = np.tile(u0, (B, 1))
    L[:, :12] = xi; U[:, :12] = xi
    return fam, xi, L, U


def run_reference_cpu(n_inst, threads, seed=1):
    """The reference's own CPU implementation of the path: vendored OSQP 0.6.2 (oracle/_ref), update_bounds +
    solve per instance, cold start, cvxpygen default settings -- timed on the host cores."""
    from oracle import ref_osqp
    if not ref_osqp.available():
        raise RuntimeError('oracle/_ref/libosqp_ref.so missing (run `make -C oracle ref` where /root/reference exists)')
    fam, xi, L, U = mpc_canonical_batch(n_inst, seed)
    r = ref_osqp.RefOSQP(fam.canon_matrix('P'), fam.canon_data('q'), fam.canon_matrix('A'),
                         fam.canon_data('l'), fam.canon_data('u'), nthreads=threads)
    out = r.solve_batch(l=L, u=U, nthreads=threads)
    return n_inst / out['seconds'], out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--batch', type=int, default=None, help='instances per GPU per step (default 100000; 50000 for portfolio_socp)')
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--cpu-sample', type=int, default=40000, help='instances in the cpu_baseline sample')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--with-grad', action='store_true', help='also time config 4: forward + backward (gradient=True)')
    ap.add_argument('--workload', default='mpc', choices=['mpc', 'portfolio_socp', 'mpc_ltv'],
                    help='mpc = the headline (BASELINE configs[1]); portfolio_socp = configs[2] through the IPM-CUDA backend; '
                         'mpc_ltv = the MPC family with per-instance matrix parameters (SURVEY row f2)')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    W = max(args.warmup, 3)
    K = args.steps
    cores = os.cpu_count() or 1
    if args.workload == 'portfolio_socp':
        args.batch = args.batch or 50000
        return main_socp(args, rank, world, local_rank, W, K, cores)
    if args.workload == 'mpc_ltv':
        args.batch = args.batch or 20000
        return main_ltv(args, rank, world, local_rank, W, K, cores)
    args.batch = args.batch or 100000

    # ------------------------------------------------------------------ reference arm (CPU, rank 0 only)
    if args.impl == 'reference':
        if rank != 0:
            return
        sample = min(args.batch, 20000)
        for _ in range(min(W, 1)):
            run_reference_cpu(2000, cores)
        t = []
        for _ in range(K):
            ips, _ = run_reference_cpu(sample, cores)
            t.append(sample / ips)
        ms = 1e3 * float(np.mean(t))
        value = sample / (ms / 1e3)
        line = {'impl': 'reference', 'metric': 'QP instances/sec (MPC n_x=12,n_u=4,N=10)', 'value': value,
                'unit': 'instances/s', 'n_gpus': args.gpus, 'steps': K, 'warmup': W, 'ms_per_step': ms,
                'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
                'config': {'workload': WORKLOAD % args.batch, 'sample': f'{sample} instances per step'},
                'cpu_baseline': {'value': value, 'unit': 'instances/s', 'cores': cores, 'kind': 'reference',
                                 'sample': f'{sample} instances/step x {K} steps, vendored OSQP 0.6.2 (oracle/_ref), {cores} host threads'},
                'e2e': {'value': value, 'unit': 'instances/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
                'gpu_launches': 0}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ B200 arm
    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), 'bench.py needs a CUDA device (no CPU fallback)'
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    from cvxpygen_b200 import standard
    mod = standard.load(FAMILY, device=local_rank).init()
    # e(multi-GPU): the only collective of the path -- one NCCL broadcast of the family constants blob
    if world > 1:
        blob = open(os.path.join(standard.code_dir(FAMILY), 'cpg_blob.bin'), 'rb').read()
        tb = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(dev)
        dist.broadcast(tb, src=0)
        mod.load_constants(bytes(tb.cpu().numpy().tobytes()))
    B = args.batch
    xi_host = np.random.default_rng(1 + rank).uniform(-1, 1, (B, 12))
    params = torch.from_numpy(xi_host).to(dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    out = None
    for _ in range(W):
        out = mod.solve_batch_device(params, out=out)
    torch.cuda.synchronize()
    launches_per_step = mod.launch_count()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank); sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    for k in range(K):
        flush.fill_(k & 0xff)                       # L2 flush, outside the event pair
        ev[k][0].record()
        out = mod.solve_batch_device(params, out=out)
        ev[k][1].record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler.active.clear()
    ms_total = sum(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / K
    value = world * B / (ms_step / 1e3)

    # ---- e2e through the public host-buffer API (pinned memory, H2D + kernel + D2H inside the timed region)
    d = mod.dims
    pin = lambda shape, dt=torch.float64: torch.empty(shape, dtype=dt).pin_memory()
    hp = pin((B, 12)); hp.copy_(torch.from_numpy(xi_host))
    hout = dict(prim=pin((B, d.n_prim)), dual=pin((B, d.n_dual)), obj=pin((B,)), pri=pin((B,)), dua=pin((B,)),
                it=pin((B,), torch.int32), st=pin((B,), torch.int32))
    e2e_steps = max(3, min(K, 5))

    def time_e2e():
        for _ in range(2):
            mod.solve_batch_pinned(hp, hout)
        if world > 1:
            dist.barrier()
        sampler.active.set()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            mod.solve_batch_pinned(hp, hout)
        t1 = time.perf_counter()
        sampler.active.clear()
        te = torch.tensor([(t1 - t0) / e2e_steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        return world * B / float(te.item())
    # default: the kernels store result rows straight into the pinned host buffers (PCIe writes overlap the solves);
    # for comparison the same call with host_zero_copy = 0 (staging buffers + D2H copies after the kernels)
    e2e_value = time_e2e()
    mod.set_solver_setting('host_zero_copy', 0)
    e2e_staged = time_e2e()
    mod.set_solver_setting('host_zero_copy', 1)
    sampler.stop_flag.set(); sampler.join()         # clocks sampled over both timed regions (device-resident and end-to-end)
    h2d = B * 12 * 8
    d2h = B * ((d.n_prim + d.n_dual) * 8 + 3 * 8 + 2 * 4)

    # ---- optional: BASELINE config 4 (MPC QP with gradient=True): backward pass on the forward solution, device-resident
    grad_info = None
    if args.with_grad:
        outg = mod.solve_batch_device(params, return_canonical=True)
        dprim = torch.randn((B, d.n_prim), dtype=torch.float64, device=dev)
        dpar = mod.gradient_batch_device(outg.sol_y, dprim)
        torch.cuda.synchronize()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(3):
            dpar = mod.gradient_batch_device(outg.sol_y, dprim, dparams=dpar)
        g1.record(); torch.cuda.synchronize()
        ms_b = g0.elapsed_time(g1) / 3
        grad_info = {'backward_ms': ms_b, 'backward_inst_per_s': B / ms_b * 1e3,
                     'forward_backward_inst_per_s': B / (ms_step + ms_b) * 1e3}

    # ---- solution quality of the last timed step (all ranks' share identical in distribution; rank 0 reports)
    st = out.status.cpu().numpy(); it = out.iter.cpu().numpy()
    frac_solved = float((st == 1).mean())
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = load_peaks()
    kernel_s = ms_step / 1e3                         # one kernel launch per step dominates: duration = step time
    achieved = (B * BYTES_PER_INSTANCE / kernel_s) / 1e9
    line = {
        'metric': 'QP instances/sec (MPC n_x=12,n_u=4,N=10)', 'value': value, 'unit': 'instances/s',
        'n_gpus': world, 'steps': K, 'warmup': W, 'ms_per_step': ms_step, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': WORKLOAD % B, 'family': FAMILY, 'batch_per_gpu': B, 'parallelism': f'batch-shard x{world}',
                   'l2': 'flushed between steps (256 MiB write), per-step CUDA events summed',
                   'mean_iter': float(it.mean()), 'frac_solved': frac_solved},
        'e2e': {'value': e2e_value, 'unit': 'instances/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                'steps': e2e_steps, 'staged_value': e2e_staged,
                'note': 'pinned host buffers through cpg_solve_batch_host, host clock: H2D of the parameters, kernels storing '
                        'prim/dual rows directly into the pinned buffers (zero-copy D2H), D2H of the info arrays; '
                        'staged_value = same call with staging buffers + D2H copies after the kernels'},
        'gpu_launches': launches_per_step * K,
        'clocks': sampler.summary(),
        'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                     'traffic': B * TRAFFIC_BYTES_PER_INSTANCE / 1e9, 'traffic_unit': 'GB per launch (ncu, profiles/r1_v7_ncu_summary.md)',
                     'peak_source': peak_src,
                     'algorithmic_bytes_per_instance': BYTES_PER_INSTANCE,
                     'note': 'on-chip design: HBM carries only parameters in / solutions out, so the HBM fraction is tiny by '
                             'construction; the binding resource is shared-memory bandwidth -- ncu: 78.2 %% of peak '
                             'shared-memory wavefronts (profiles/r1_v7_ncu_summary.md); fp64 fraction = %.4f of 37 TFLOP/s '
                             'nominal' % (value / world * FLOP_PER_INSTANCE / 37e12)},
    }
    if grad_info:
        line['config']['gradient'] = grad_info
    if not args.no_cpu_baseline:
        try:
            ips, _ = run_reference_cpu(args.cpu_sample, cores)
            line['cpu_baseline'] = {'value': ips, 'unit': 'instances/s', 'cores': cores, 'kind': 'reference',
                                    'sample': f'{args.cpu_sample} instances of the same workload, vendored OSQP 0.6.2 '
                                              f'(oracle/_ref), {cores} host threads, update_bounds+solve per instance'}
        except Exception as e:      # the checker is test infrastructure: report, do not fail the GPU number
            line['cpu_baseline'] = {'value': None, 'unit': 'instances/s', 'cores': cores, 'kind': 'reference', 'sample': f'unavailable: {e}'}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------------------------
# BASELINE.json configs[2]: portfolio SOCP through the IPM-CUDA backend (python bench.py --workload portfolio_socp)
def socp_params(B, seed):
    rng = np.random.default_rng(seed)
    return np.ascontiguousarray(np.c_[rng.standard_normal((B, 100)), np.abs(1 / 100 + 0.01 * rng.standard_normal((B, 100)))])


def run_reference_cpu_socp(n_inst, threads, seed=3):
    """The reference's CPU path for this family: vendored ECOS 2.0.8 (oracle/_ref), ECOS_updateData + ECOS_solve per
    instance, one workspace per host thread."""
    import concurrent.futures as cf
    from cvxpygen_b200 import families
    from oracle import ref_ecos
    if not ref_ecos.available():
        raise RuntimeError('oracle/_ref/libecos_ref.so missing (run `make -C oracle ref` where /root/reference exists)')
    fam = families.portfolio_socp(100, 10)
    P = socp_params(n_inst, seed)
    c0, b0 = fam.canon_data('c'), fam.canon_data('b')
    Cb = np.tile(c0, (n_inst, 1)); Cb[:, :100] = -P[:, :100]
    Bb = np.tile(b0, (n_inst, 1)); Bb[:, 11:111] = -P[:, 100:]
    nw = max(1, min(threads, n_inst))
    refs = [ref_ecos.RefECOS(c0, fam.canon_matrix('A'), b0, fam.canon_matrix('G'), fam.canon_data('h'), 601, [12, 102]) for _ in range(nw)]

    def work(k):
        sl = slice(k * n_inst // nw, (k + 1) * n_inst // nw)
        return refs[k].solve_batch(c=Cb[sl], b=Bb[sl])
    t0 = time.perf_counter()
    with cf.ThreadPoolExecutor(nw) as ex:
        list(ex.map(work, range(nw)))
    return n_inst / (time.perf_counter() - t0), nw


def main_socp(args, rank, world, local_rank, W, K, cores):
    metric = 'SOCP instances/sec (portfolio n=100 assets)'
    B = args.batch
    if args.impl == 'reference':
        if rank != 0:
            return
        sample = min(B, 4096)
        run_reference_cpu_socp(512, cores)
        t = []
        for _ in range(K):
            ips, nw = run_reference_cpu_socp(sample, cores)
            t.append(sample / ips)
        ms = 1e3 * float(np.mean(t)); value = sample / (ms / 1e3)
        print(json.dumps({'impl': 'reference', 'metric': metric, 'value': value, 'unit': 'instances/s', 'n_gpus': args.gpus, 'steps': K,
                          'warmup': W, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
                          'data': 'synthetic', 'config': {'workload': SOCP_WORKLOAD % B, 'sample': f'{sample} instances per step'},
                          'cpu_baseline': {'value': value, 'unit': 'instances/s', 'cores': nw, 'kind': 'reference',
                                           'sample': f'{sample} instances/step x {K} steps, vendored ECOS 2.0.8 (oracle/_ref), {nw} host threads'},
                          'e2e': {'value': value, 'unit': 'instances/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}, 'gpu_launches': 0}))
        return
    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), 'bench.py needs a CUDA device (no CPU fallback)'
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    from cvxpygen_b200 import standard
    mod = standard.load(SOCP_FAMILY, device=local_rank).init()
    P_host = socp_params(B, 3 + rank)
    params = torch.from_numpy(P_host).to(dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    out = None
    for _ in range(W):
        out = mod.solve_batch_device(params, out=out)
    torch.cuda.synchronize()
    launches_per_step = mod.launch_count()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank); sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    for k in range(K):
        flush.fill_(k & 0xff)
        ev[k][0].record()
        out = mod.solve_batch_device(params, out=out)
        ev[k][1].record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler.active.clear()
    t = torch.tensor([sum(a.elapsed_time(b) for a, b in ev)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / K
    value = world * B / (ms_step / 1e3)
    d = mod.dims
    pin = lambda shape, dt=torch.float64: torch.empty(shape, dtype=dt).pin_memory()
    hp = pin((B, d.n_param)); hp.copy_(torch.from_numpy(P_host))
    hout = dict(prim=pin((B, d.n_prim)), dual=pin((B, d.n_dual)), obj=pin((B,)), pri=pin((B,)), dua=pin((B,)),
                it=pin((B,), torch.int32), st=pin((B,), torch.int32))
    mod.solve_batch_pinned(hp, hout)
    e2e_steps = 2
    sampler.active.set()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        mod.solve_batch_pinned(hp, hout)
    te = torch.tensor([(time.perf_counter() - t0) / e2e_steps], dtype=torch.float64, device=dev)
    sampler.stop_flag.set(); sampler.join()         # clocks sampled over both timed regions
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    st = out.status.cpu().numpy(); it = out.iter.cpu().numpy()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = load_peaks()
    achieved = B * SOCP_BYTES_PER_INSTANCE / (ms_step / 1e3) / 1e9
    line = {'metric': metric, 'value': value, 'unit': 'instances/s', 'n_gpus': world, 'steps': K, 'warmup': W, 'ms_per_step': ms_step,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': SOCP_WORKLOAD % B, 'family': SOCP_FAMILY, 'batch_per_gpu': B, 'parallelism': f'batch-shard x{world}',
                       'l2': 'flushed between steps (256 MiB write), per-step CUDA events summed',
                       'mean_iter': float(it.mean()), 'frac_optimal': float((st == 0).mean()), 'threads_per_cta': int(d.threads_per_cta),
                       'smem_bytes_per_cta': int(d.smem_bytes)},
            'e2e': {'value': world * B / float(te.item()), 'unit': 'instances/s', 'h2d_bytes_per_step': B * d.n_param * 8,
                    'd2h_bytes_per_step': B * ((d.n_prim + d.n_dual) * 8 + 32), 'steps': e2e_steps,
                    'note': 'pinned host buffers through cpg_socp_solve_batch_host: H2D, kernel, D2H; host clock'},
            'gpu_launches': launches_per_step * K, 'clocks': sampler.summary(),
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                         'traffic': B * SOCP_TRAFFIC_BYTES_PER_INSTANCE / 1e9,
                         'traffic_unit': 'GB per launch, scaled from the ncu capture at batch 1184 (profiles/r1_ipm_v7_ncu_summary.md: 2.37 MB read + '
                                         '0.20 MB written -- the result rows of so small a batch stay in L2; at batch 50000 they are written back: + 2.6 KB/instance)',
                         'peak_source': peak_src, 'algorithmic_bytes_per_instance': SOCP_BYTES_PER_INSTANCE,
                         'note': 'one CTA per instance with the whole interior-point state (iterate, scalings, numeric LDL\' factor, work '
                                 'vectors: ~215 KB) in shared memory; HBM carries parameters in / solutions out only; the binding resource '
                                 'is issue latency inside barrier-separated sparse phases (profiles/r1_ipm_v4_ncu_summary.md)'}}
    if not args.no_cpu_baseline:
        try:
            n = min(args.cpu_sample, 4096)
            ips, nw = run_reference_cpu_socp(n, cores)
            line['cpu_baseline'] = {'value': ips, 'unit': 'instances/s', 'cores': nw, 'kind': 'reference',
                                    'sample': f'{n} instances of the same workload, vendored ECOS 2.0.8 (oracle/_ref), {nw} host threads'}
        except Exception as e:
            line['cpu_baseline'] = {'value': None, 'unit': 'instances/s', 'cores': cores, 'kind': 'reference', 'sample': f'unavailable: {e}'}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------------------------
# MPC QP with per-instance matrix parameters (python bench.py --workload mpc_ltv); the reference arm is
# osqp_update_P_A + update_lin_cost/bounds + osqp_solve per instance on the host cores (oracle/_ref).
LTV_FAMILY = 'mpc_ltv_12_4_10'
LTV_WORKLOAD = ('MPC QP (n_x=12,n_u=4,N=10) with per-instance dynamics A, B and stage costs (220-entry parameter row; dense-pattern '
                'A: nnz 2092) batch=%d per GPU, ADMM-CUDA matrix-parameter kernel, OSQP default settings')
LTV_BYTES_PER_INSTANCE = 220 * 8 + (172 + 172) * 8 + 40
LTV_TRAFFIC_BYTES_PER_INSTANCE = 3001      # ncu: (38.08 MB + 21.94 MB) / 20000 instances


def ltv_canonical(B, seed):
    from cvxpygen_b200 import families
    fam = families.mpc_ltv(12, 4, 10)
    params = families.mpc_ltv_batch(fam, B, seed=seed)
    th = np.tile(fam.theta_default(), (B, 1))
    for pn, v in params.items():
        p = fam.param(pn)
        th[:, p.col:p.col + p.size] = v
    Px = np.asarray((fam.maps['P'] @ th.T).T); Ax = np.asarray((fam.maps['A'] @ th.T).T)
    l = np.clip(np.asarray(th @ fam.maps['l'].T.toarray()), -1e30, 1e30)
    u = np.clip(np.asarray(th @ fam.maps['u'].T.toarray()), -1e30, 1e30)
    return fam, params, Px, Ax, l, u


def run_reference_cpu_ltv(n_inst, threads, seed=31):
    from oracle import ref_osqp
    if not ref_osqp.available():
        raise RuntimeError('oracle/_ref/libosqp_ref.so missing (run `make -C oracle ref` where /root/reference exists)')
    fam, _, Px, Ax, l, u = ltv_canonical(n_inst, seed)
    r = ref_osqp.RefOSQP(fam.canon_matrix('P'), fam.canon_data('q'), fam.canon_matrix('A'),
                         fam.canon_data('l'), fam.canon_data('u'), nthreads=threads)
    out = r.solve_batch_mat(Px=Px, Ax=Ax, l=l, u=u, nthreads=threads)
    return n_inst / out['seconds'], out


def main_ltv(args, rank, world, local_rank, W, K, cores):
    metric = 'QP instances/sec (MPC n_x=12,n_u=4,N=10, per-instance matrices)'
    B = args.batch
    if args.impl == 'reference':
        if rank != 0:
            return
        sample = min(B, 4000)
        run_reference_cpu_ltv(256, cores)
        t = []
        for _ in range(K):
            ips, _ = run_reference_cpu_ltv(sample, cores)
            t.append(sample / ips)
        ms = 1e3 * float(np.mean(t)); value = sample / (ms / 1e3)
        print(json.dumps({'impl': 'reference', 'metric': metric, 'value': value, 'unit': 'instances/s', 'n_gpus': args.gpus, 'steps': K,
                          'warmup': W, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
                          'data': 'synthetic', 'config': {'workload': LTV_WORKLOAD % B, 'sample': f'{sample} instances per step'},
                          'cpu_baseline': {'value': value, 'unit': 'instances/s', 'cores': cores, 'kind': 'reference',
                                           'sample': f'{sample} instances/step x {K} steps, vendored OSQP 0.6.2 (oracle/_ref): '
                                                     f'osqp_update_P_A + update_bounds + osqp_solve per instance, {cores} host threads'},
                          'e2e': {'value': value, 'unit': 'instances/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}, 'gpu_launches': 0}))
        return
    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), 'bench.py needs a CUDA device (no CPU fallback)'
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    from cvxpygen_b200 import standard, families
    mod = standard.load(LTV_FAMILY, device=local_rank).init()
    fam = families.mpc_ltv(12, 4, 10)
    P_host = mod.pack_params(families.mpc_ltv_batch(fam, B, seed=31 + rank))
    params = torch.from_numpy(P_host).to(dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    out = None
    for _ in range(W):
        out = mod.solve_batch_device(params, out=out)
    torch.cuda.synchronize()
    launches_per_step = mod.launch_count()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank); sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    for k in range(K):
        flush.fill_(k & 0xff)
        ev[k][0].record()
        out = mod.solve_batch_device(params, out=out)
        ev[k][1].record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler.active.clear()
    t = torch.tensor([sum(a.elapsed_time(b) for a, b in ev)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / K
    value = world * B / (ms_step / 1e3)
    d = mod.dims
    pin = lambda shape, dt=torch.float64: torch.empty(shape, dtype=dt).pin_memory()
    hp = pin((B, d.n_param)); hp.copy_(torch.from_numpy(P_host))
    hout = dict(prim=pin((B, d.n_prim)), dual=pin((B, d.n_dual)), obj=pin((B,)), pri=pin((B,)), dua=pin((B,)),
                it=pin((B,), torch.int32), st=pin((B,), torch.int32))
    mod.solve_batch_pinned(hp, hout)
    e2e_steps = 3
    sampler.active.set()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        mod.solve_batch_pinned(hp, hout)
    te = torch.tensor([(time.perf_counter() - t0) / e2e_steps], dtype=torch.float64, device=dev)
    sampler.stop_flag.set(); sampler.join()         # clocks sampled over both timed regions
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    st = out.status.cpu().numpy(); it = out.iter.cpu().numpy()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = load_peaks()
    achieved = B * LTV_BYTES_PER_INSTANCE / (ms_step / 1e3) / 1e9
    line = {'metric': metric, 'value': value, 'unit': 'instances/s', 'n_gpus': world, 'steps': K, 'warmup': W, 'ms_per_step': ms_step,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': LTV_WORKLOAD % B, 'family': LTV_FAMILY, 'batch_per_gpu': B, 'parallelism': f'batch-shard x{world}',
                       'l2': 'flushed between steps (256 MiB write), per-step CUDA events summed',
                       'mean_iter': float(it.mean()), 'frac_solved': float((st == 1).mean())},
            'e2e': {'value': world * B / float(te.item()), 'unit': 'instances/s', 'h2d_bytes_per_step': B * d.n_param * 8,
                    'd2h_bytes_per_step': B * ((d.n_prim + d.n_dual) * 8 + 32), 'steps': e2e_steps,
                    'note': 'pinned host buffers through cpg_solve_batch_host: H2D of the parameter rows, kernel (zero-copy result rows), D2H of the info arrays; host clock'},
            'gpu_launches': launches_per_step * K, 'clocks': sampler.summary(),
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                         'traffic': B * LTV_TRAFFIC_BYTES_PER_INSTANCE / 1e9,
                         'traffic_unit': 'GB per launch (ncu dram__bytes_read+write at batch 20000, profiles/r1_matpar_v1_ncu_summary.md: '
                                         '38.1 MB read + 21.9 MB written; part of the result rows is still in the 126 MB L2 when the kernel ends)',
                         'peak_source': peak_src, 'algorithmic_bytes_per_instance': LTV_BYTES_PER_INSTANCE,
                         'note': 'one warp per instance: equilibration, KKT assembly, numeric LDL\' and the ADMM loop all on chip '
                                 '(factor in shared memory, tables in L2); HBM carries the parameter row in and the solution rows out; '
                                 'the binding resource is the dependent chain of the per-instance triangular solves (profiles/r1_matpar_ncu_summary.md)'}}
    if not args.no_cpu_baseline:
        try:
            n = min(args.cpu_sample, 4000)
            ips, _ = run_reference_cpu_ltv(n, cores)
            line['cpu_baseline'] = {'value': ips, 'unit': 'instances/s', 'cores': cores, 'kind': 'reference',
                                    'sample': f'{n} instances of the same workload, vendored OSQP 0.6.2 (oracle/_ref): osqp_update_P_A + '
                                              f'update_bounds + osqp_solve per instance, {cores} host threads'}
        except Exception as e:
            line['cpu_baseline'] = {'value': None, 'unit': 'instances/s', 'cores': cores, 'kind': 'reference', 'sample': f'unavailable: {e}'}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
