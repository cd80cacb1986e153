#!/usr/bin/env python
"""bench.py -- headline benchmark of the batched-solve hot path (BASELINE.json `metric`).

    python bench.py --gpus N --steps K --warmup W                    # this repo's sm_100a backend
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU path (oracle/_ref)

Headline (the JSON line's top level): a *step* is one batched solve of `--batch` (default 100 000) MPC QP instances
(n_x=12, n_u=4, N=10; BASELINE.json configs[1]) per GPU with synthetic x_init ~ U[-1,1]^12 (seed 1 + rank).  Weak scaling:
every rank solves its own batch; value = all instances / max-over-ranks device time.

The same line carries a `workloads` block with the other BASELINE configs measured the same way in the same run --
`portfolio_socp` (configs[2], 50 000 instances, IPM-CUDA), `mpc_grad` (configs[3]: forward + backward, gradient=True),
`mpc_ltv` (per-instance matrix parameters) -- each with value / ms_per_step / roofline / cpu_baseline /
e2e / parity; `--workload X` makes X the headline instead, `--no-workloads` drops the block.

`parity`: after the timed region the results of a sample of the SAME batch are compared with the compiled unmodified
reference (oracle/_ref: vendored OSQP 0.6.2 / ECOS 2.0.8 / the reference's generated gradient C) and max relative errors
plus the fraction of identical iteration counts are printed (rank 0 only).

Timing: W >= 3 warm-up steps; every timed step is bracketed by CUDA events on the launching stream; between steps a
256 MiB buffer is written to flush the 126 MB L2 (outside the event pairs); the K event times are summed; max over ranks.
`roofline.achieved` divides the algorithmic bytes of one launch by the dominant kernel's own duration, measured live by
CUDA events the library records around that launch (cpg_b200_kernel_times).  `roofline.traffic` / `roofline.binding`
are read from profiles/r2_kernel_metrics.json (written by profiles/summarize_ncu.py from the ncu capture of this command).
`e2e` is the same metric through the public host-buffer API (pinned host memory: H2D of the parameters, kernels, D2H of
every result array) timed with the host clock around the synchronous call.
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

FLOP_PER_MPC_INSTANCE = 0.5e6
KERNEL_METRICS = os.path.join(ROOT, 'profiles', 'r2_kernel_metrics.json')


def load_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


def kernel_metrics(kernel):
    """ncu-derived numbers of the committed capture of this binary: dram bytes per instance and the binding resource."""
    try:
        with open(KERNEL_METRICS) as f:
            return json.load(f).get(kernel)
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons sampled DURING the timed regions."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.rows = index, threading.Event(), []
        self.active = threading.Event()           # samples are taken only while a timed region is running

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


def relmax_rows(a, b):
    """max over instances of ||a_i - b_i|| / ||b_i||"""
    nb = np.maximum(np.linalg.norm(b, axis=1), 1e-12)
    return float((np.linalg.norm(a - b, axis=1) / nb).max())


# =====================================================================================================================
# Workloads.  Each knows: its family, its synthetic parameter batch, its algorithmic bytes per instance, the reference's
# CPU implementation of the same path (oracle/_ref) and how to compare a sample with it.
class Workload:
    key = family = metric = kernel = None
    default_batch = 100000
    cpu_sample = 40000         # instances of the cpu_baseline sample
    ref_sample = None          # instances per step of the reference arm (None = the full batch)
    parity_sample = 10000
    sub_steps = 5              # timed steps when run inside the `workloads` block

    def describe(self, B): raise NotImplementedError
    def host_params(self, B, seed): raise NotImplementedError
    def bytes_per_instance(self, d): raise NotImplementedError
    def run_reference(self, n, threads, seed): raise NotImplementedError      # -> instances/s
    def parity(self, mod, P_host): raise NotImplementedError
    def reference_note(self, cores): raise NotImplementedError
    note = ''

    # one step on device-resident inputs (asynchronous on torch's current stream)
    def alloc(self, mod, params):
        return {'out': None}

    def step(self, mod, params, st):
        st['out'] = mod.solve_batch_device(params, out=st['out'])

    def dominant_ms(self, mod):
        return mod.kernel_times()['main']

    def launches_per_step(self, mod):
        return mod.launch_count()                    # kernels launched by the last solve call

    def quality(self, st):
        it = st['out'].iter.cpu().numpy(); s = st['out'].status.cpu().numpy()
        return {'mean_iter': float(it.mean()), 'frac_solved': float((s == self.ok_status).mean())}
    ok_status = 1

    def pinned_io(self, mod, B, P_host):
        import torch
        d = mod.dims
        pin = lambda shape, dt=torch.float64: torch.empty(shape, dtype=dt).pin_memory()
        hp = pin((B, d.n_param)); hp.copy_(torch.from_numpy(P_host))
        hout = dict(prim=pin((B, d.n_prim)), dual=pin((B, d.n_dual)), obj=pin((B,)), pri=pin((B,)), dua=pin((B,)),
                    it=pin((B,), torch.int32), st=pin((B,), torch.int32))
        return hp, hout

    def e2e_step(self, mod, hp, hout):
        mod.solve_batch_pinned(hp, hout)

    def e2e_bytes(self, d, B):
        return B * d.n_param * 8, B * ((d.n_prim + d.n_dual) * 8 + 3 * 8 + 2 * 4)


class MpcWorkload(Workload):
    key, family, kernel = 'mpc', 'mpc_12_4_10', 'admm_multi_kernel'
    metric = 'QP instances/sec (MPC n_x=12,n_u=4,N=10)'
    default_batch, parity_sample = 100000, 100000        # the whole batch is compared (1.3 s of host time)
    note = ('on-chip design: HBM carries only parameters in / solutions out, so the HBM fraction is tiny by construction; '
            'see `binding` for the resource that limits the kernel')

    def describe(self, B):
        return ('MPC QP (n_x=12,n_u=4,N=10) batch=%d per GPU, ADMM-CUDA backend, OSQP default settings (eps 1e-3, adaptive rho)' % B)

    def host_params(self, B, seed):
        return np.random.default_rng(seed).uniform(-1, 1, (B, 12))

    def bytes_per_instance(self, d):
        return 12 * 8 + (172 + 172) * 8 + 40            # 96 B in + 2.75 KB out + 40 B info

    def canonical(self, P_host):
        from cvxpygen_b200 import families
        fam = families.mpc(12, 4, 10)
        B = P_host.shape[0]
        l0, u0 = fam.canon_data('l'), fam.canon_data('u')
        L = np.tile(l0, (B, 1)); U = np.tile(u0, (B, 1))
        L[:, :12] = P_host; U[:, :12] = P_host
        return fam, L, U

    def reference(self, P_host, threads):
        """vendored OSQP 0.6.2 (oracle/_ref): osqp_update_bounds + osqp_solve per instance, cold start, default settings"""
        from oracle import ref_osqp
        if not ref_osqp.available():
            raise RuntimeError('oracle/_ref/libosqp_ref.so missing (run `make -C oracle ref` where /root/reference exists)')
        fam, L, U = self.canonical(P_host)
        r = ref_osqp.RefOSQP(fam.canon_matrix('P'), fam.canon_data('q'), fam.canon_matrix('A'),
                             fam.canon_data('l'), fam.canon_data('u'), nthreads=threads)
        return r.solve_batch(l=L, u=U, nthreads=threads)

    def run_reference(self, n, threads, seed):
        out = self.reference(self.host_params(n, seed), threads)
        return n / out['seconds']

    def reference_note(self, cores):
        return f'vendored OSQP 0.6.2 (oracle/_ref), {cores} host threads, osqp_update_bounds + osqp_solve per instance'

    def parity(self, mod, P_host):
        res = mod.solve_batch(P_host, return_canonical=True)
        ora = self.reference(P_host, os.cpu_count() or 1)
        ok = np.isin(ora['status'], [1, 2, -2])
        return {'sample': int(P_host.shape[0]), 'vs': 'oracle/_ref: unmodified vendored OSQP 0.6.2',
                'max_rel_x': relmax_rows(res.sol_x[ok], ora['x'][ok]), 'max_rel_y': relmax_rows(res.sol_y[ok], ora['y'][ok]),
                'iter_equal_frac': float((res.cpg_info.iter == ora['iter']).mean()),
                'status_equal_frac': float((res.cpg_info.status == ora['status']).mean()),
                'max_abs_obj': float(np.abs(res.cpg_info.obj_val[ok] - ora['obj'][ok]).max())}


class LtvWorkload(MpcWorkload):
    key, family, kernel = 'mpc_ltv', 'mpc_ltv_12_4_10', 'admm_matpar_kernel'
    metric = 'QP instances/sec (MPC n_x=12,n_u=4,N=10, per-instance matrices)'
    default_batch, cpu_sample, ref_sample, parity_sample = 20000, 4000, 4000, 2000
    note = ('one warp per instance: equilibration, KKT assembly, numeric LDL\' and the ADMM loop all on chip (factor in shared '
            'memory, tables in L2); HBM carries the parameter row in and the solution rows out')

    def describe(self, B):
        return ('MPC QP (n_x=12,n_u=4,N=10) with per-instance dynamics A, B and stage costs (220-entry parameter row; dense-pattern '
                'A: nnz 2092) batch=%d per GPU, ADMM-CUDA matrix-parameter kernel, OSQP default settings' % B)

    def fam(self):
        from cvxpygen_b200 import families
        return families.mpc_ltv(12, 4, 10)

    def host_params(self, B, seed):
        from cvxpygen_b200 import families
        fam = self.fam()
        params = families.mpc_ltv_batch(fam, B, seed=30 + seed)
        cols = [np.asarray(params[n]).reshape(B, -1) for n in ('A', 'B', 'qdiag', 'rdiag', 'x_init')]
        return np.ascontiguousarray(np.concatenate(cols, axis=1))

    def bytes_per_instance(self, d):
        return 220 * 8 + (172 + 172) * 8 + 40

    def reference(self, P_host, threads):
        from oracle import ref_osqp
        if not ref_osqp.available():
            raise RuntimeError('oracle/_ref/libosqp_ref.so missing (run `make -C oracle ref` where /root/reference exists)')
        fam = self.fam()
        B = P_host.shape[0]
        th = np.tile(fam.theta_default(), (B, 1))
        col = 0
        for n in ('A', 'B', 'qdiag', 'rdiag', 'x_init'):
            p = fam.param(n)
            th[:, p.col:p.col + p.size] = P_host[:, col:col + p.size]; col += p.size
        Px = np.asarray((fam.maps['P'] @ th.T).T); Ax = np.asarray((fam.maps['A'] @ th.T).T)
        l = np.clip(np.asarray(th @ fam.maps['l'].T.toarray()), -1e30, 1e30)
        u = np.clip(np.asarray(th @ fam.maps['u'].T.toarray()), -1e30, 1e30)
        r = ref_osqp.RefOSQP(fam.canon_matrix('P'), fam.canon_data('q'), fam.canon_matrix('A'),
                             fam.canon_data('l'), fam.canon_data('u'), nthreads=threads)
        return r.solve_batch_mat(Px=Px, Ax=Ax, l=l, u=u, nthreads=threads)

    def reference_note(self, cores):
        return (f'vendored OSQP 0.6.2 (oracle/_ref), {cores} host threads, osqp_update_P_A + osqp_update_bounds + osqp_solve '
                'per instance')


class GradWorkload(MpcWorkload):
    """BASELINE configs[3]: MPC QP with gradient=True -- one step = forward solve + backward pass."""
    key, family, kernel = 'mpc_grad', 'mpc_12_4_10', 'qp_grad_kernel'
    metric = 'QP instances/sec forward+backward (MPC n_x=12,n_u=4,N=10, gradient=True)'
    default_batch, cpu_sample, ref_sample, parity_sample = 100000, 16000, 16000, 10000
    note = ('roofline block = the backward kernel (qp_grad_kernel: per-instance numeric LDL\' of the active-set KKT system + 4 '
            'solves, one warp per instance); the forward kernel is the headline workload\'s')

    def describe(self, B):
        return ('MPC QP (n_x=12,n_u=4,N=10) with gradient=True: forward solve + diff-through-KKT backward pass, batch=%d per GPU, '
                'ADMM-CUDA backend' % B)

    def bytes_per_instance(self, d):        # backward kernel: sol_y + dprim in, dparams out
        return (172 + 172 + 12) * 8

    def alloc(self, mod, params):
        import torch
        B = params.shape[0]
        g = torch.Generator(device=params.device).manual_seed(5)
        return {'out': None, 'dprim': torch.randn((B, mod.dims.n_prim), dtype=torch.float64, device=params.device, generator=g),
                'dpar': None}

    def step(self, mod, params, st):
        st['out'] = mod.solve_batch_device(params, out=st['out'], return_canonical=True)
        st['dpar'] = mod.gradient_batch_device(st['out'].sol_y, st['dprim'], dparams=st['dpar'])

    def dominant_ms(self, mod):
        return mod.kernel_times()['grad']

    def launches_per_step(self, mod):
        return 2 + mod.launch_count()                # forward: main + tail kernel; the last call was the backward kernel

    def pinned_io(self, mod, B, P_host):
        import torch
        hp, hout = super().pinned_io(mod, B, P_host)
        d = mod.dims
        pin = lambda shape: torch.empty(shape, dtype=torch.float64).pin_memory()
        hout['sol_y'] = pin((B, d.n_con))
        hout['dprim'] = pin((B, d.n_prim)); hout['dprim'].copy_(torch.randn((B, d.n_prim), dtype=torch.float64,
                                                                           generator=torch.Generator().manual_seed(5)))
        hout['dparams'] = pin((B, d.n_param))
        return hp, hout

    def e2e_step(self, mod, hp, hout):
        mod.solve_batch_pinned(hp, hout)
        mod.gradient_batch_pinned(hout['sol_y'], hout['dprim'], hout['dparams'])

    def e2e_bytes(self, d, B):
        h2d, d2h = super().e2e_bytes(d, B)
        return h2d + B * (d.n_con + d.n_prim) * 8, d2h + B * (d.n_con + d.n_param) * 8

    @staticmethod
    def grad_libs(k):
        """k private copies of the reference's generated gradient C (its workspace is one static struct per library)"""
        import shutil, tempfile
        src = os.path.join(ROOT, 'oracle', '_ref', 'libgrad_ref_mpc_12_4_10.so')
        if not os.path.exists(src):
            raise RuntimeError('oracle/_ref/libgrad_ref_mpc_12_4_10.so missing (python oracle/build_grad_ref.py mpc_12_4_10)')
        td = tempfile.mkdtemp(prefix='cpg_gradref_')
        out = []
        for i in range(k):
            dst = os.path.join(td, f'libgrad_ref_{i}.so'); shutil.copyfile(src, dst); out.append(dst)
        return out

    def reference_backward(self, fam, x, y, dprim, threads):
        import concurrent.futures as cf
        from oracle.build_grad_ref import grad_ref_batch
        B = x.shape[0]
        prim_idx = np.concatenate([v.indices for v in fam.variables])
        dx = np.zeros((B, fam.n_var)); dx[:, prim_idx] = dprim
        nw = max(1, min(threads, B // 64 or 1))
        libs = self.grad_libs(nw)
        m = fam.n_eq + fam.n_ineq
        sl = [slice(k * B // nw, (k + 1) * B // nw) for k in range(nw)]
        t0 = time.perf_counter()
        with cf.ThreadPoolExecutor(nw) as ex:
            parts = list(ex.map(lambda k: grad_ref_batch(libs[k], fam.n_var, m, x[sl[k]], y[sl[k]], dx[sl[k]]), range(nw)))
        sec = time.perf_counter() - t0
        dq, dl, du = (np.concatenate([p[i] for p in parts]) for i in range(3))
        # un-canonicalisation: d x_init = Ml' dl + Mu' du (cvxpygen/writer.py:268-303); x_init enters l and u rows 0..11
        return dl[:, :12] + du[:, :12], sec

    def run_reference(self, n, threads, seed):
        P_host = self.host_params(n, seed)
        fam, L, U = self.canonical(P_host)
        fwd = self.reference(P_host, threads)
        dprim = np.random.default_rng(5).standard_normal((n, 172))
        _, sec_b = self.reference_backward(fam, fwd['x'], fwd['y'], dprim, threads)
        return n / (fwd['seconds'] + sec_b)

    def reference_note(self, cores):
        return (f'vendored OSQP 0.6.2 forward + the reference\'s own generated gradient C (cpg_osqp_grad_compute.c rendered from its '
                f'templates, oracle/build_grad_ref.py) backward, {cores} host threads (one static gradient workspace per thread)')

    def parity(self, mod, P_host):
        B = P_host.shape[0]
        res = mod.solve_batch(P_host, return_canonical=True)
        dprim = np.random.default_rng(5).standard_normal((B, mod.dims.n_prim))
        got = mod.gradient_batch(res.sol_y, dprim)['x_init']
        fam, _, _ = self.canonical(P_host)
        ref, _ = self.reference_backward(fam, res.sol_x, res.sol_y, dprim, os.cpu_count() or 1)
        return {'sample': int(B), 'vs': 'oracle/_ref: the reference\'s generated cpg_osqp_gradient C on the same forward solution',
                'max_rel_dparams': relmax_rows(got, ref),
                'max_rel_dparams_batch': float(np.abs(got - ref).max() / np.abs(ref).max())}


class SocpWorkload(Workload):
    key, family, kernel = 'portfolio_socp', 'portfolio_socp_100_10', 'ipm_kernel'
    metric = 'SOCP instances/sec (portfolio n=100 assets)'
    default_batch, cpu_sample, ref_sample, parity_sample, sub_steps = 50000, 4096, 4096, 4096, 3
    ok_status = 0
    note = ('one CTA per instance with the whole interior-point state (iterate, scalings, numeric LDL\' factor, work vectors) in '
            'shared memory; HBM carries parameters in / solutions out only')

    def describe(self, B):
        return ('portfolio SOCP (n=100 assets, 10 factors; 512 vars, 111 eq, 715 cone rows) batch=%d per GPU, IPM-CUDA backend, '
                'ECOS default settings (tol 1e-8)' % B)

    def host_params(self, B, seed):
        rng = np.random.default_rng(2 + seed)
        return np.ascontiguousarray(np.c_[rng.standard_normal((B, 100)), np.abs(1 / 100 + 0.01 * rng.standard_normal((B, 100)))])

    def bytes_per_instance(self, d):
        return 200 * 8 + (210 + 112) * 8 + 40            # a, w_prev in; w, delta_w, f + duals out; info (~4.2 KB)

    def quality(self, st):
        q = super().quality(st)
        q['frac_optimal'] = q.pop('frac_solved')
        return q

    def reference(self, P_host, threads):
        """vendored ECOS 2.0.8 (oracle/_ref): ECOS_updateData + ECOS_solve per instance, one workspace per host thread"""
        import concurrent.futures as cf
        from cvxpygen_b200 import families
        from oracle import ref_ecos
        if not ref_ecos.available():
            raise RuntimeError('oracle/_ref/libecos_ref.so missing (run `make -C oracle ref` where /root/reference exists)')
        fam = families.portfolio_socp(100, 10)
        n = P_host.shape[0]
        c0, b0 = fam.canon_data('c'), fam.canon_data('b')
        Cb = np.tile(c0, (n, 1)); Cb[:, :100] = -P_host[:, :100]
        Bb = np.tile(b0, (n, 1)); Bb[:, 11:111] = -P_host[:, 100:]
        nw = max(1, min(threads, n))
        refs = [ref_ecos.RefECOS(c0, fam.canon_matrix('A'), b0, fam.canon_matrix('G'), fam.canon_data('h'), 601, [12, 102])
                for _ in range(nw)]
        sl = [slice(k * n // nw, (k + 1) * n // nw) for k in range(nw)]
        t0 = time.perf_counter()
        with cf.ThreadPoolExecutor(nw) as ex:
            parts = list(ex.map(lambda k: refs[k].solve_batch(c=Cb[sl[k]], b=Bb[sl[k]]), range(nw)))
        sec = time.perf_counter() - t0
        out = {k: np.concatenate([p[k] for p in parts]) for k in ('x', 'y', 'z', 's', 'iter', 'exitflag', 'pcost')}
        out['seconds'] = sec
        return fam, out

    def run_reference(self, n, threads, seed):
        _, out = self.reference(self.host_params(n, seed), threads)
        return n / out['seconds']

    def reference_note(self, cores):
        return f'vendored ECOS 2.0.8 (oracle/_ref), {cores} host threads, ECOS_updateData + ECOS_solve per instance'

    def parity(self, mod, P_host):
        res = mod.solve_batch(P_host, return_canonical=True)
        fam, ora = self.reference(P_host, os.cpu_count() or 1)
        prim_ref = np.concatenate([ora['x'][:, v.indices] for v in fam.variables], axis=1)
        dual_ref = np.concatenate([ora[d.vec][:, d.indices] for d in fam.duals], axis=1)
        return {'sample': int(P_host.shape[0]), 'vs': 'oracle/_ref: unmodified vendored ECOS 2.0.8',
                'max_rel_prim': relmax_rows(res.prim, prim_ref), 'max_rel_dual': relmax_rows(res.dual, dual_ref),
                'max_rel_x': relmax_rows(res.sol_x, ora['x']), 'max_rel_z': relmax_rows(res.sol_z, ora['z']),
                'iter_equal_frac': float((res.cpg_info.iter == ora['iter']).mean()),
                'status_equal_frac': float((res.cpg_info.status == ora['exitflag']).mean())}


class SocpMatWorkload(SocpWorkload):
    """Config 3's problem with the reference example's MATRIX parameters batched too: F (in A) and d_sqrt (in G) per instance."""
    key, family, kernel = 'portfolio_socp_mat', 'portfolio_socp_mat_100_10', 'ipm_kernel'
    metric = 'SOCP instances/sec (portfolio n=100 assets, per-instance F and d_sqrt)'
    default_batch, cpu_sample, ref_sample, parity_sample, sub_steps = 20000, 2048, 2048, 2048, 3
    note = ('ipm_kernel with IPM_MATPAR: per instance the G / A entries are canonicalised, re-equilibrated (ECOS set_equilibration) and '
            'written into the KKT image inside the kernel; 10.4 KB of parameters per instance in')

    def describe(self, B):
        return ('portfolio SOCP (n=100 assets, 10 factors) with per-instance factor loadings F (100x10, in A) and d_sqrt (in G), '
                'batch=%d per GPU, IPM-CUDA backend, ECOS default settings (tol 1e-8)' % B)

    def _fam(self):
        from cvxpygen_b200 import families
        return families.portfolio_socp(100, 10, matrix_params=True)

    def host_params(self, B, seed):
        fam = self._fam()
        rng = np.random.default_rng(5 + seed)
        a = rng.standard_normal((B, 100))
        wp = np.abs(1 / 100 + 0.01 * rng.standard_normal((B, 100)))
        F = fam.param('F').default[None, :] + 0.25 * rng.standard_normal((B, 1000))
        d = fam.param('d_sqrt').default[None, :] * rng.uniform(0.5, 1.5, (B, 100))
        return np.ascontiguousarray(np.c_[a, wp, F, d])

    def bytes_per_instance(self, d):
        return 1300 * 8 + (210 + 112) * 8 + 40

    def reference(self, P_host, threads):
        """vendored ECOS 2.0.8: raw G / A / c / b / h of every instance through ECOS_updateData (re-equilibration) + ECOS_solve"""
        import concurrent.futures as cf
        from oracle import ref_ecos
        if not ref_ecos.available():
            raise RuntimeError('oracle/_ref/libecos_ref.so missing (run `make -C oracle ref` where /root/reference exists)')
        fam = self._fam()
        n = P_host.shape[0]
        th = np.tile(fam.theta_default(), (n, 1))
        col = 0
        for nm in ('a', 'w_prev', 'F', 'd_sqrt'):
            p_ = fam.param(nm)
            th[:, p_.col:p_.col + p_.size] = P_host[:, col:col + p_.size]; col += p_.size
        data = {k: np.asarray((fam.maps[k] @ th.T).T) for k in ('c', 'b', 'h', 'A', 'G')}
        nw = max(1, min(threads, n))
        refs = [ref_ecos.RefECOS(fam.canon_data('c'), fam.canon_matrix('A'), fam.canon_data('b'), fam.canon_matrix('G'),
                                 fam.canon_data('h'), 601, [12, 102]) for _ in range(nw)]
        sl = [slice(k * n // nw, (k + 1) * n // nw) for k in range(nw)]
        t0 = time.perf_counter()
        with cf.ThreadPoolExecutor(nw) as ex:
            parts = list(ex.map(lambda k: refs[k].solve_batch(c=data['c'][sl[k]], b=data['b'][sl[k]], h=data['h'][sl[k]],
                                                              G=data['G'][sl[k]], A=data['A'][sl[k]]), range(nw)))
        sec = time.perf_counter() - t0
        out = {k: np.concatenate([p_[k] for p_ in parts]) for k in ('x', 'y', 'z', 's', 'iter', 'exitflag', 'pcost')}
        out['seconds'] = sec
        return fam, out

    def reference_note(self, cores):
        return f'vendored ECOS 2.0.8 (oracle/_ref), {cores} host threads, per-instance G / A values: ECOS_updateData (re-equilibration) + ECOS_solve'


WORKLOADS = {w.key: w for w in (MpcWorkload(), SocpWorkload(), GradWorkload(), LtvWorkload(), SocpMatWorkload())}


# =====================================================================================================================
def measure(wl, B, W, K, rank, world, local_rank, dev, sampler, flush, with_cpu, with_parity, cores, cpu_sample=None):
    """One workload on this rank's GPU; returns the record (rank 0) or None."""
    import torch
    import torch.distributed as dist
    from cvxpygen_b200 import standard
    mod = standard.load(wl.family, device=local_rank).init()
    P_host = wl.host_params(B, 1 + rank)
    params = torch.from_numpy(P_host).to(dev)
    st = wl.alloc(mod, params)
    for _ in range(W):
        wl.step(mod, params, st)
    torch.cuda.synchronize()
    launches_per_step = wl.launches_per_step(mod)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    kms = []
    sampler.active.set()
    for k in range(K):
        flush.fill_(k & 0xff)                       # L2 flush, outside the event pair
        ev[k][0].record()
        wl.step(mod, params, st)
        ev[k][1].record()
        if k == K - 1:
            kms.append(wl.dominant_ms(mod))         # the library's own events around the dominant kernel (last step)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler.active.clear()
    t = torch.tensor([sum(a.elapsed_time(b) for a, b in ev)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / K
    value = world * B / (ms_step / 1e3)

    # ---- e2e through the public host-buffer API (pinned memory, H2D + kernels + D2H inside the timed region)
    hp, hout = wl.pinned_io(mod, B, P_host)
    e2e_steps = max(2, min(K, 5))

    def time_e2e():
        for _ in range(2):
            wl.e2e_step(mod, hp, hout)
        if world > 1:
            dist.barrier()
        sampler.active.set()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            wl.e2e_step(mod, hp, hout)
        t1 = time.perf_counter()
        sampler.active.clear()
        te = torch.tensor([(t1 - t0) / e2e_steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        return world * B / float(te.item())
    e2e_value = time_e2e()
    e2e_staged = None
    if wl.key == 'mpc':          # for comparison: staging buffers + D2H copies after the kernels instead of zero-copy rows
        mod.set_solver_setting('host_zero_copy', 0)
        e2e_staged = time_e2e()
        mod.set_solver_setting('host_zero_copy', 1)
    quality = wl.quality(st)
    if rank != 0:
        return None
    d = mod.dims
    h2d, d2h = wl.e2e_bytes(d, B)
    peak, peak_src = load_peaks()
    bpi = wl.bytes_per_instance(d)
    k_ms = kms[-1] if kms and kms[-1] else ms_step
    achieved = (B * bpi / (k_ms / 1e3)) / 1e9
    km = kernel_metrics(wl.kernel) or {}
    rec = {
        'metric': wl.metric, 'value': value, 'unit': 'instances/s', 'n_gpus': world, 'steps': K, 'warmup': W,
        'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': dict({'workload': wl.describe(B), 'family': wl.family, 'batch_per_gpu': B, 'parallelism': f'batch-shard x{world}',
                        'l2': 'flushed between steps (256 MiB write), per-step CUDA events summed'}, **quality),
        'e2e': {'value': e2e_value, 'unit': 'instances/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h, 'steps': e2e_steps,
                'note': 'pinned host buffers through the C-ABI host entry, host clock: H2D of the parameters, kernels (result rows '
                        'stored directly into the pinned buffers where the entry supports it), D2H of the remaining arrays'},
        'gpu_launches': launches_per_step * K,
        'roofline': {'bound': 'hbm', 'kernel': wl.kernel, 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                     'kernel_ms': k_ms, 'kernel_ms_source': 'CUDA events recorded by the library around the launch (last timed step)',
                     'traffic': (B * km['dram_bytes_per_instance'] / 1e9) if km.get('dram_bytes_per_instance') else None,
                     'traffic_unit': 'GB per launch = ncu dram__bytes_read+write per instance x batch',
                     'traffic_source': km.get('source'), 'binding': km.get('binding'),
                     'peak_source': peak_src, 'algorithmic_bytes_per_instance': bpi, 'note': wl.note},
    }
    if e2e_staged is not None:
        rec['e2e']['staged_value'] = e2e_staged
    if wl.key == 'mpc':
        rec['roofline']['fp64_frac'] = value / world * FLOP_PER_MPC_INSTANCE / 37e12
    if with_parity:
        try:
            rec['parity'] = wl.parity(mod, P_host[:min(B, wl.parity_sample)])
        except Exception as e:
            rec['parity'] = {'unavailable': f'{type(e).__name__}: {e}'}
    if with_cpu:
        n = min(cpu_sample or wl.cpu_sample, wl.cpu_sample if cpu_sample is None else cpu_sample)
        try:
            ips = wl.run_reference(n, cores, 1)
            rec['cpu_baseline'] = {'value': ips, 'unit': 'instances/s', 'cores': cores, 'kind': 'reference',
                                   'sample': f'{n} instances of the same workload, ' + wl.reference_note(cores)}
        except Exception as e:      # the checker is test infrastructure: report, do not fail the GPU number
            rec['cpu_baseline'] = {'value': None, 'unit': 'instances/s', 'cores': cores, 'kind': 'reference', 'sample': f'unavailable: {e}'}
    return rec


def reference_record(wl, B, W, K, cores, n_gpus, bounded):
    sample = B if (wl.ref_sample is None and not bounded) else min(B, wl.ref_sample or wl.cpu_sample)
    wl.run_reference(min(sample, 2000), cores, 1)      # warm-up (library load, page-in)
    t = []
    for _ in range(K):
        t.append(sample / wl.run_reference(sample, cores, 1))
    ms = 1e3 * float(np.mean(t))
    value = sample / (ms / 1e3)
    return {'impl': 'reference', 'metric': wl.metric, 'value': value, 'unit': 'instances/s', 'n_gpus': n_gpus, 'steps': K, 'warmup': W,
            'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': wl.describe(B), 'family': wl.family, 'batch_per_gpu': B},
            'cpu_baseline': {'value': value, 'unit': 'instances/s', 'cores': cores, 'kind': 'reference',
                             'sample': f'{sample} instances/step x {K} steps, ' + wl.reference_note(cores)},
            'e2e': {'value': value, 'unit': 'instances/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--batch', type=int, default=None, help='instances per GPU per step of the headline workload')
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--cpu-sample', type=int, default=None, help='instances in the cpu_baseline sample of the headline workload')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-parity', action='store_true')
    ap.add_argument('--no-workloads', action='store_true', help='skip the `workloads` block (the other BASELINE configs)')
    ap.add_argument('--with-grad', action='store_true', help='(kept for compatibility: the gradient workload is part of `workloads`)')
    ap.add_argument('--workload', default='mpc', choices=list(WORKLOADS),
                    help='headline workload: mpc = BASELINE configs[1]; portfolio_socp = configs[2]; mpc_grad = configs[3]; '
                         'mpc_ltv = per-instance matrix parameters (SURVEY row f2)')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    W = max(args.warmup, 3)
    K = args.steps
    cores = os.cpu_count() or 1
    head = WORKLOADS[args.workload]
    B = args.batch or head.default_batch
    others = [] if (args.no_workloads or args.workload != 'mpc') else [WORKLOADS[k] for k in ('portfolio_socp', 'mpc_grad', 'mpc_ltv', 'portfolio_socp_mat')]

    # ------------------------------------------------------------------ reference arm (CPU, rank 0 only)
    if args.impl == 'reference':
        if rank != 0:
            return
        line = reference_record(head, B, W, K, cores, args.gpus, bounded=False)
        if others:
            line['workloads'] = {}
            for wl in others:
                try:
                    r = reference_record(wl, wl.default_batch, 1, max(1, min(K, 2)), cores, args.gpus, bounded=True)
                    line['workloads'][wl.key] = {k: r[k] for k in ('metric', 'value', 'unit', 'ms_per_step', 'config', 'cpu_baseline')}
                except Exception as e:
                    line['workloads'][wl.key] = {'unavailable': f'{type(e).__name__}: {e}'}
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
    # e(multi-GPU): the only collective of the path -- one NCCL broadcast of the family constants blob
    if world > 1 and not head.key.startswith('portfolio_socp'):
        mod = standard.load(head.family, device=local_rank).init()
        blob = open(os.path.join(standard.code_dir(head.family), 'cpg_blob.bin'), 'rb').read()
        tb = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(dev)
        dist.broadcast(tb, src=0)
        mod.load_constants(bytes(tb.cpu().numpy().tobytes()))
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    sampler = ClockSampler(local_rank); sampler.start()
    line = measure(head, B, W, K, rank, world, local_rank, dev, sampler, flush, with_cpu=not args.no_cpu_baseline,
                   with_parity=not args.no_parity, cores=cores, cpu_sample=args.cpu_sample)
    subs = {}
    for wl in others:
        try:
            r = measure(wl, wl.default_batch, 3, max(2, min(K, wl.sub_steps)), rank, world, local_rank, dev, sampler, flush,
                        with_cpu=not args.no_cpu_baseline, with_parity=not args.no_parity, cores=cores)
        except Exception as e:
            if world > 1:
                raise
            r = {'unavailable': f'{type(e).__name__}: {e}'}
        if rank == 0:
            subs[wl.key] = r
    sampler.stop_flag.set(); sampler.join()         # clocks sampled over every timed region of this run
    if rank == 0:
        line['clocks'] = sampler.summary()
        if subs:
            line['workloads'] = subs
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
