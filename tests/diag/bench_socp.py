#!/usr/bin/env python
"""Throughput of the IPM-CUDA backend on BASELINE config 3 (portfolio SOCP n=100, batch 50k), device-resident, plus the
compiled reference (ECOS 2.0.8) on the host cores for a bounded sample.  Prints one JSON line."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=50000)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=1)
    ap.add_argument('--cpu-sample', type=int, default=2048)
    ap.add_argument('--code-dir', default=None, help='a generated IPM-CUDA directory (default: the standard portfolio family)')
    args = ap.parse_args()
    import torch
    from cvxpygen_b200 import standard, families
    if args.code_dir:
        from cvxpygen_b200 import runtime
        m = runtime.load(args.code_dir)
    else:
        m = standard.load('portfolio_socp_100_10')
    B = args.batch
    rng = np.random.default_rng(3)
    a = rng.standard_normal((B, 100)); wp = np.abs(1 / 100 + 0.01 * rng.standard_normal((B, 100)))
    P = torch.from_numpy(np.ascontiguousarray(np.c_[a, wp])).cuda()
    out = None
    for _ in range(args.warmup):
        out = m.solve_batch_device(P, out=out)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    ev[0].record()
    for i in range(args.steps):
        out = m.solve_batch_device(P, out=out)
        ev[i + 1].record()
    torch.cuda.synchronize()
    ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    it = out.iter.cpu().numpy(); st = out.status.cpu().numpy()
    line = dict(metric='portfolio SOCP instances/sec (n=100 assets), IPM-CUDA', value=B / (np.mean(ms) * 1e-3), unit='instances/s',
                batch=B, ms_per_step=float(np.mean(ms)), ms_all=ms, iters_mean=float(it.mean()), solved_frac=float((st == 0).mean()),
                threads_per_cta=int(m.dims.threads_per_cta), smem_bytes=int(m.dims.smem_bytes))
    if args.cpu_sample:
        from oracle import ref_ecos
        if ref_ecos.available():
            fam = families.portfolio_socp()
            c0, b0 = fam.canon_data('c'), fam.canon_data('b')
            n = args.cpu_sample
            Cb = np.tile(c0, (n, 1)); Cb[:, :100] = -a[:n]
            Bb = np.tile(b0, (n, 1)); Bb[:, 11:111] = -wp[:n]
            import concurrent.futures as cf
            cores = os.cpu_count() or 1
            nw = min(cores, 64)

            def work(k):
                r = ref_ecos.RefECOS(c0, fam.canon_matrix('A'), b0, fam.canon_matrix('G'), fam.canon_data('h'), 601, [12, 102])
                sl = slice(k * n // nw, (k + 1) * n // nw)
                return r.solve_batch(c=Cb[sl], b=Bb[sl])['seconds']
            t0 = time.perf_counter()
            with cf.ThreadPoolExecutor(nw) as ex:
                secs = list(ex.map(work, range(nw)))
            wall = time.perf_counter() - t0
            line['cpu_baseline'] = dict(value=n / wall, unit='instances/s', cores=nw, kind='reference',
                                        sample=f'{n} instances, ECOS 2.0.8 compiled from the reference tree, {nw} threads',
                                        per_core=n / sum(secs))
    print(json.dumps(line))


if __name__ == '__main__':
    main()
