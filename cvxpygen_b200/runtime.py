"""ctypes binding of a generated ADMM-CUDA solver library (libcpg_b200.so).

This file is copied verbatim into every generated code directory as ``cpg_module.py``; it plays
the role of the reference's pybind11 module ``cpg_module`` (emitter cvxpygen/utils.py:1331-1412,
structs TPL/cpg_module.hpp.jinja2:10-76):

  reference                               here
  --------------------------------------  ------------------------------------------------------
  cpg_module.solve(upd, par) -> result    Module.solve(upd, par) -> result      (batch of one)
  cpg_module.set_solver_default_settings  Module.set_solver_default_settings()
  cpg_module.set_solver_<name>(v)         Module.set_solver_<name>(v)   (AttributeError if unknown)
  <prefix>cpg_params / cpg_updated / ...  Module.cpg_params() / cpg_updated() ... plain attribute bags
  (none)                                  Module.solve_batch(params)            NEW: host arrays
  (none)                                  Module.solve_batch_device(tensor)     NEW: torch CUDA tensors

There is no CPU fallback: if the library or a CUDA device is missing, construction raises.
"""
import ctypes as C
import json
import os
import time
from types import SimpleNamespace

import numpy as np

STATUS_STRINGS = {1: 'solved', 2: 'solved inaccurate', 3: 'primal infeasible inaccurate',
                  4: 'dual infeasible inaccurate', -2: 'maximum iterations reached',
                  -3: 'primal infeasible', -4: 'dual infeasible', -7: 'problem non convex',
                  -10: 'unsolved', -100: 'handed off'}   # osqp_sources/src/auxil.c:655-679


class CpgB200Settings(C.Structure):
    _fields_ = [('max_iter', C.c_int), ('check_termination', C.c_int), ('scaled_termination', C.c_int),
                ('warm_start', C.c_int), ('adaptive_rho', C.c_int), ('adaptive_rho_interval', C.c_int),
                ('scaling', C.c_int), ('host_zero_copy', C.c_int),
                ('eps_abs', C.c_double), ('eps_rel', C.c_double), ('eps_prim_inf', C.c_double),
                ('eps_dual_inf', C.c_double), ('alpha', C.c_double), ('adaptive_rho_tolerance', C.c_double)]


class CpgB200Dims(C.Structure):
    _fields_ = [('n_var', C.c_int), ('n_con', C.c_int), ('n_param', C.c_int), ('n_prim', C.c_int),
                ('n_dual', C.c_int), ('blob_bytes', C.c_int), ('warps_per_cta', C.c_int), ('smem_bytes', C.c_int)]


# settings the reference exposes for OSQP and their cvxpy aliases (cvxpygen/solvers/osqp.py:102-115)
SETTING_ALIASES = {'warm_starting': 'warm_start'}
READONLY_SETTINGS = ('scaling',)


class Module:
    """One loaded solver library bound to one CUDA device."""

    def __init__(self, code_dir, device=0):
        self.code_dir = os.path.abspath(code_dir)
        with open(os.path.join(self.code_dir, 'cpg_meta.json')) as f:
            self.meta = json.load(f)
        self.prefix = self.meta['prefix']
        path = os.path.join(self.code_dir, 'libcpg_b200.so')
        if not os.path.exists(path):
            raise RuntimeError(f'{path} is missing: run cvxpygen_b200.codegen.compile_code (nvcc, sm_100a) first; '
                               'there is no CPU fallback')
        self.lib = C.CDLL(path)
        self._fn('cpg_b200_last_error').restype = C.c_char_p
        self.dims = CpgB200Dims()
        self._check(self._fn('cpg_b200_dims')(C.byref(self.dims)))
        self.settings = CpgB200Settings()
        self.set_solver_default_settings()
        self.device = device
        self._initialised = False

    # ---- plumbing
    def _fn(self, name):
        return getattr(self.lib, self.prefix + name)

    @staticmethod
    def _expect(a, shape, what):
        """The C entries take raw pointers: a mismatched array would be read past its end.  None in `shape` = any."""
        got = tuple(a.shape)
        if len(got) != len(shape) or any(s is not None and s != g for s, g in zip(shape, got)):
            raise ValueError(f'{what}: expected shape {tuple("B" if s is None else s for s in shape)}, got {got}')

    def _check(self, rc):
        if rc != 0:
            msg = self._fn('cpg_b200_last_error')()
            raise RuntimeError(f'cpg_b200 error {rc}: {msg.decode() if msg else ""}')

    def init(self):
        if not self._initialised:
            self._check(self._fn('cpg_b200_init')(C.c_int(self.device)))
            self._initialised = True
        return self

    def load_constants(self, blob: bytes):
        self.init()
        self._check(self._fn('cpg_b200_load_constants')(C.c_char_p(blob), C.c_int(len(blob))))

    def update_shared_params(self, values):
        """Change user parameters that are shared by the whole batch (e.g. a matrix parameter).  Role of the reference's
        osqp_update_data_mat branch of cpg_solve (cvxpygen/solvers/osqp.py:20-33 -> re-scale + refactor): the offline
        setup is re-run on the host and every constants table is re-uploaded.  Needs the cvxpygen_b200 package."""
        import pickle
        import numpy as _np
        from cvxpygen_b200.offline.qp_setup import setup_qp_family
        with open(os.path.join(self.code_dir, 'cpg_family.pkl'), 'rb') as f:
            saved = pickle.load(f)
        fam = saved['family']
        if not hasattr(self, '_theta'):
            self._theta = fam.theta_default()
        for name, val in values.items():
            p = fam.param(name)                                   # AttributeError for unknown names
            if name in saved['batch_params']:
                raise ValueError(f'{name} is a batched parameter: pass it per instance to solve_batch')
            v = _np.asarray(val, dtype=float)
            v = v.flatten(order='F') if v.ndim > 1 and v.size == int(_np.prod(p.shape)) and p.size == v.size else v.ravel()
            if v.size != p.size:
                raise ValueError(f'parameter {name} stores {p.size} entries, got {v.size}')
            self._theta[p.col:p.col + p.size] = v
        st = setup_qp_family(fam, saved['batch_params'], theta=self._theta, rho=saved['rho'], sigma=saved['sigma'],
                             scaling=saved['scaling'], max_group_rows=saved.get('max_group_rows', 32),
                             allow_trailing=saved.get('allow_trailing', True))
        if st.solve_source != saved['solve_source']:
            raise RuntimeError('the new parameter values change the sparsity structure of the KKT factor schedule: '
                               'regenerate the code (cpg.generate_code)')
        self.init()
        gS0 = _np.asarray(st.grad_S0, dtype='<f8').tobytes()
        b = lambda x: (C.c_char_p(x), C.c_int(len(x)))
        self._check(self._fn('cpg_b200_load_constants_all')(*b(st.blob), *b(st.blob_compact), *b(st.tail_blob),
                                                           *b(st.grad_blob), *b(gS0)))
        if getattr(st, 'dmma_blob', b''):      # tensor-core main kernel: its coefficient tables changed with the factor
            self._check(self._fn('cpg_b200_load_dmma_constants')(*b(st.dmma_blob)))
        if st.mat_blob:       # matrix-parameter family: base values / maps of the P and A entries changed as well
            self._check(self._fn('cpg_b200_load_mat_constants')(*b(st.mat_blob)))
        return st

    def launch_count(self):
        return int(self._fn('cpg_b200_launch_count')())

    def kernel_times(self):
        """Device milliseconds of the kernels of the last solve / gradient call (CUDA events recorded by the library on the
        caller's stream around each launch): dict(main=, tail=, grad=), None where nothing was launched."""
        t = [C.c_float(-1.0) for _ in range(3)]
        self._check(self._fn('cpg_b200_kernel_times')(*[C.byref(v) for v in t]))
        return {k: (None if v.value < 0 else float(v.value)) for k, v in zip(('main', 'tail', 'grad'), t)}

    # ---- settings (b2: cpg_set_solver_default_settings / cpg_set_solver_<name>)
    def set_solver_default_settings(self):
        self._fn('cpg_b200_default_settings')(C.byref(self.settings))

    def set_solver_setting(self, name, value):
        name = SETTING_ALIASES.get(name, name)
        if name in READONLY_SETTINGS or name not in dict(CpgB200Settings._fields_):
            raise AttributeError(f'Solver setting "{name}" not available.')   # TPL/cpg_solver.py.jinja2:59-60
        setattr(self.settings, name, value)

    def __getattr__(self, item):
        if item.startswith('set_solver_'):
            name = item[len('set_solver_'):]
            return lambda v: self.set_solver_setting(name, v)
        raise AttributeError(item)

    # ---- attribute bags mirroring the pybind classes
    def cpg_params(self):
        return SimpleNamespace(**{p['name']: list(p['default']) if p['size'] > 1 else p['default'][0]
                                  for p in self.meta['params']})

    def cpg_updated(self):
        return SimpleNamespace(**{p['name']: False for p in self.meta['params']})

    # gradient=True flavour of the pybind module (cvxpygen/utils.py:1272-1328, 1364-1405): cpg_gsol / cpg_vdelta / gradient
    def cpg_gsol(self):
        return SimpleNamespace(primal=None, dual=None)

    def cpg_vdelta(self):
        return SimpleNamespace(**{v['name']: ([0.0] * v['size'] if v['size'] > 1 else 0.0) for v in self.meta['variables']})

    def gradient(self, vdelta, gsol, use_sol):
        """cpg_module.gradient(vdelta, gsol, use_sol) -> pdelta: the backward pass of ONE instance (a batch of one on the GPU)
        at the canonical solution in `gsol` (use_sol) or at the last `solve`.  pdelta carries one attribute per user
        parameter; parameters that are shared by the batch in this generated code (folded into the constants) get None."""
        if use_sol:
            sx, sy = np.asarray(gsol.primal, dtype=np.float64)[None, :], np.asarray(gsol.dual, dtype=np.float64)[None, :]
        else:
            if getattr(self, '_last', None) is None:
                raise RuntimeError('cpg_module.gradient: no previous solve and no solution passed (use_sol)')
            sx, sy = self._last['sol_x'], self._last['sol_y']
        d = np.zeros((1, self.dims.n_prim))
        for v in self.meta['variables']:
            d[0, v['offset']:v['offset'] + v['size']] = np.atleast_1d(np.asarray(getattr(vdelta, v['name']), dtype=np.float64)).ravel()
        if self.has_matrix_params:
            if getattr(self, '_last', None) is None:
                raise RuntimeError('cpg_module.gradient: a family with per-instance matrices needs the parameters of the last solve')
            res = self.gradient_batch_mat(self._last['params'], sx, sy, d)
        else:
            res = self.gradient_batch(sy, d)
        out = {}
        for p in self.meta['params']:
            if p['batched']:
                out[p['name']] = res[p['name']][0].tolist() if p['size'] > 1 else float(res[p['name']][0, 0])
            else:
                out[p['name']] = None
        return SimpleNamespace(**out)

    # ---- packing helpers
    def pack_params(self, params, B=None):
        """dict name -> (B, size) | (B, *shape) | (size,) broadcast  ->  (B, n_param) float64, batched params only.
        Matrices are flattened in Fortran order like the reference (TPL/cpg_solver.py.jinja2:26-34)."""
        bp = [p for p in self.meta['params'] if p['batched']]
        unknown = set(params) - {p['name'] for p in self.meta['params']}
        if unknown:
            raise AttributeError(f'{sorted(unknown)[0]} is not a parameter.')
        shared_given = [n for n in params if n not in {p['name'] for p in bp}]
        if shared_given:
            raise ValueError(f'parameters {shared_given} are shared by the whole batch in this generated code; '
                             'update them with update_shared_params (host re-setup), not per instance')
        flat = {}
        for p in bp:
            if p['name'] not in params:
                continue
            a = np.asarray(params[p['name']], dtype=np.float64)
            shp, sz = tuple(p['shape']), p['size']
            if a.shape == shp or a.shape == (sz,) or a.size == 1 == sz and a.ndim == 0:
                a = a.flatten(order='F').reshape(1, sz)                 # one value, broadcast over the batch
            elif a.shape[1:] == shp and len(shp) > 1:
                a = a.transpose([0] + list(range(a.ndim - 1, 0, -1))).reshape(a.shape[0], sz)   # per-instance F-order
            elif a.ndim == 2 and a.shape[1] == sz:
                pass
            elif a.ndim == 1 and sz == 1:
                a = a.reshape(-1, 1)
            else:
                raise ValueError(f"parameter {p['name']}: cannot interpret shape {a.shape} for size {sz}")
            flat[p['name']] = a
            if a.shape[0] > 1:
                if B is not None and B != a.shape[0]:
                    raise ValueError('inconsistent batch sizes')
                B = a.shape[0]
        B = 1 if B is None else B
        out = np.empty((B, self.dims.n_param), dtype=np.float64)
        col = 0
        for p in bp:
            sz = p['size']
            a = flat.get(p['name'], np.asarray(p['default'], dtype=np.float64).reshape(1, sz))
            out[:, col:col + sz] = np.broadcast_to(a, (B, sz))
            col += sz
        return out

    def unpack(self, prim, dual):
        res_p, res_d = {}, {}
        for v in self.meta['variables']:
            a = prim[:, v['offset']:v['offset'] + v['size']]
            res_p[v['name']] = a.reshape((a.shape[0],) + tuple(v['shape']), order='F') if len(v['shape']) > 1 else a
        for d in self.meta['duals']:
            a = dual[:, d['offset']:d['offset'] + d['size']]
            res_d[d['name']] = a
        return res_p, res_d

    # ---- NEW: batched solve, host arrays (goes through cpg_solve_batch_host: H2D + kernel + D2H)
    def solve_batch(self, params, x0=None, y0=None, return_canonical=False, **settings):
        self.init()
        for k, v in settings.items():
            self.set_solver_setting(k, v)
        P = params if isinstance(params, np.ndarray) else self.pack_params(params)
        P = np.ascontiguousarray(P, dtype=np.float64)
        d = self.dims
        self._expect(P, (None, d.n_param), 'params')
        B = P.shape[0]
        prim = np.empty((B, d.n_prim)); dual = np.empty((B, d.n_dual))
        solx = np.empty((B, d.n_var)) if return_canonical else None
        soly = np.empty((B, d.n_con)) if return_canonical else None
        obj = np.empty(B); pri = np.empty(B); dua = np.empty(B)
        it = np.empty(B, dtype=np.int32); st = np.empty(B, dtype=np.int32)
        s = self.settings
        if x0 is not None and y0 is not None:
            x0 = np.ascontiguousarray(x0, dtype=np.float64); y0 = np.ascontiguousarray(y0, dtype=np.float64)
            self._expect(x0, (B, d.n_var), 'x0'); self._expect(y0, (B, d.n_con), 'y0')
            s = CpgB200Settings.from_buffer_copy(bytes(self.settings)); s.warm_start = 1

        def p(a, t=C.c_double):
            return None if a is None else a.ctypes.data_as(C.POINTER(t))
        t0 = time.perf_counter()
        self._check(self._fn('cpg_solve_batch_host')(C.c_int(B), p(P), p(x0), p(y0), p(prim), p(dual), p(solx), p(soly),
                                                     p(obj), p(it, C.c_int), p(st, C.c_int), p(pri), p(dua), C.byref(s)))
        t1 = time.perf_counter()
        pr, du = self.unpack(prim, dual)
        info = SimpleNamespace(obj_val=obj, iter=it, status=st, pri_res=pri, dua_res=dua, time=t1 - t0)
        return SimpleNamespace(cpg_prim=pr, cpg_dual=du, cpg_info=info, prim=prim, dual=dual, sol_x=solx, sol_y=soly)

    def solve_batch_multi(self, params, devices=None, x0=None, y0=None, return_canonical=False, out=None, **settings):
        """The batch on SEVERAL devices of this node from one call (cpg_solve_batch_host_multi): contiguous shards, one host
        thread per device inside the library, nothing exchanged between devices.  devices: list of CUDA device indices
        (default: every visible device).  Same result object as solve_batch.  out: optional dict of caller-owned float64 arrays
        `prim` (B, n_prim), `dual` (B, n_dual), `obj`, `pri`, `dua` (B) and int32 `it`, `st` (B) -- numpy views of PINNED buffers
        (e.g. torch.empty(..., pin_memory=True).numpy()) let every device store its result rows straight into host memory."""
        for k, v in settings.items():
            self.set_solver_setting(k, v)
        if devices is None:
            import torch
            devices = list(range(torch.cuda.device_count()))
        if not devices:
            raise RuntimeError('no CUDA device visible (there is no CPU fallback)')
        P = params if isinstance(params, np.ndarray) else self.pack_params(params)
        P = np.ascontiguousarray(P, dtype=np.float64)
        d = self.dims
        self._expect(P, (None, d.n_param), 'params')
        B = P.shape[0]
        if out is not None:
            prim, dual, obj, pri, dua, it, st = (out[k] for k in ('prim', 'dual', 'obj', 'pri', 'dua', 'it', 'st'))
            self._expect(prim, (B, d.n_prim), 'out[prim]'); self._expect(dual, (B, d.n_dual), 'out[dual]')
            for a, nm in ((obj, 'obj'), (pri, 'pri'), (dua, 'dua'), (it, 'it'), (st, 'st')):
                if a.shape != (B,) or not a.flags.c_contiguous or a.dtype != (np.int32 if nm in ('it', 'st') else np.float64):
                    raise ValueError(f'out[{nm}]: expected a contiguous ({B},) array of the right dtype')
        else:
            prim = np.empty((B, d.n_prim)); dual = np.empty((B, d.n_dual))
            obj = np.empty(B); pri = np.empty(B); dua = np.empty(B)
            it = np.empty(B, dtype=np.int32); st = np.empty(B, dtype=np.int32)
        solx = np.empty((B, d.n_var)) if return_canonical else None
        soly = np.empty((B, d.n_con)) if return_canonical else None
        s = self.settings
        if x0 is not None and y0 is not None:
            x0 = np.ascontiguousarray(x0, dtype=np.float64); y0 = np.ascontiguousarray(y0, dtype=np.float64)
            self._expect(x0, (B, d.n_var), 'x0'); self._expect(y0, (B, d.n_con), 'y0')
            s = CpgB200Settings.from_buffer_copy(bytes(self.settings)); s.warm_start = 1
        dev = (C.c_int * len(devices))(*devices)

        def p(a, t=C.c_double):
            return None if a is None else a.ctypes.data_as(C.POINTER(t))
        t0 = time.perf_counter()
        self._check(self._fn('cpg_solve_batch_host_multi')(C.c_int(len(devices)), dev, C.c_int(B), p(P), p(x0), p(y0), p(prim), p(dual),
                                                           p(solx), p(soly), p(obj), p(it, C.c_int), p(st, C.c_int), p(pri), p(dua),
                                                           C.byref(s)))
        t1 = time.perf_counter()
        pr, du = self.unpack(prim, dual)
        info = SimpleNamespace(obj_val=obj, iter=it, status=st, pri_res=pri, dua_res=dua, time=t1 - t0)
        return SimpleNamespace(cpg_prim=pr, cpg_dual=du, cpg_info=info, prim=prim, dual=dual, sol_x=solx, sol_y=soly)

    def solve_batch_pinned(self, params, out, x0=None, y0=None):
        """Host-buffer entry on caller-owned (ideally pinned) torch CPU tensors: no allocation, no copies on the
        Python side.  out: dict with prim, dual, obj, pri, dua (float64) and it, st (int32) tensors."""
        self.init()
        B = params.shape[0]
        ptr = lambda t: C.c_void_p(0 if t is None else t.data_ptr())
        s = self.settings
        if x0 is not None and y0 is not None:
            s = CpgB200Settings.from_buffer_copy(bytes(self.settings)); s.warm_start = 1
        self._check(self._fn('cpg_solve_batch_host')(C.c_int(B), ptr(params), ptr(x0), ptr(y0), ptr(out.get('prim')),
                                                     ptr(out.get('dual')), ptr(out.get('sol_x')), ptr(out.get('sol_y')),
                                                     ptr(out['obj']), ptr(out['it']), ptr(out['st']), ptr(out['pri']),
                                                     ptr(out['dua']), C.byref(s)))
        return out

    # ---- NEW: batched solve on torch CUDA tensors already resident in HBM (asynchronous on the current stream)
    def solve_batch_device(self, params, x0=None, y0=None, out=None, return_canonical=False):
        import torch
        self.init()
        assert params.is_cuda and params.dtype == torch.float64 and params.is_contiguous()
        d = self.dims
        self._expect(params, (None, d.n_param), 'params')
        B = params.shape[0]
        dev = params.device
        if dev.index is not None and dev.index != self.device:
            raise ValueError(f'params live on cuda:{dev.index}, this library is bound to cuda:{self.device}')
        for nm, t, w in (('x0', x0, d.n_var), ('y0', y0, d.n_con)):
            if t is not None:
                self._expect(t, (B, w), nm)
        if out is None:
            out = SimpleNamespace(
                prim=torch.empty((B, d.n_prim), dtype=torch.float64, device=dev),
                dual=torch.empty((B, d.n_dual), dtype=torch.float64, device=dev),
                sol_x=torch.empty((B, d.n_var), dtype=torch.float64, device=dev) if return_canonical else None,
                sol_y=torch.empty((B, d.n_con), dtype=torch.float64, device=dev) if return_canonical else None,
                obj_val=torch.empty(B, dtype=torch.float64, device=dev),
                pri_res=torch.empty(B, dtype=torch.float64, device=dev),
                dua_res=torch.empty(B, dtype=torch.float64, device=dev),
                iter=torch.empty(B, dtype=torch.int32, device=dev),
                status=torch.empty(B, dtype=torch.int32, device=dev))
        s = self.settings
        if x0 is not None and y0 is not None:
            s = CpgB200Settings.from_buffer_copy(bytes(self.settings)); s.warm_start = 1
        ptr = lambda t: C.c_void_p(0 if t is None else t.data_ptr())
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        self._check(self._fn('cpg_solve_batch_device')(C.c_int(B), ptr(params), ptr(x0), ptr(y0), ptr(out.prim), ptr(out.dual),
                                                       ptr(out.sol_x), ptr(out.sol_y), ptr(out.obj_val), ptr(out.iter),
                                                       ptr(out.status), ptr(out.pri_res), ptr(out.dua_res), C.byref(s), stream))
        return out

    # ---- NEW: batched backward pass (gradient=True).  dprim: dict name -> (B, *shape) or packed (B, n_prim) array
    def pack_dprim(self, dvars, B):
        d = np.zeros((B, self.dims.n_prim))
        for v in self.meta['variables']:
            if v['name'] in dvars:
                a = np.asarray(dvars[v['name']], dtype=np.float64)
                a = a.reshape(B, -1, order='F') if a.ndim <= 2 else np.stack([t.flatten(order='F') for t in a])
                d[:, v['offset']:v['offset'] + v['size']] = a
        return d

    def unpack_dparams(self, dparams):
        out, col = {}, 0
        for p in self.meta['params']:
            if p['batched']:
                out[p['name']] = dparams[:, col:col + p['size']]
                col += p['size']
        return out

    @property
    def has_matrix_params(self):
        return bool(self.meta.get('matrix_params'))

    def gradient_batch_mat(self, params, sol_x, sol_y, dprim, return_canonical=False):
        """Backward pass of a family with per-instance matrix parameters (host arrays): `params` as given to solve_batch
        (dict or packed rows), sol_x / sol_y the canonical solution of the forward pass.  Returns the dict of parameter
        gradients[, dq, dl, du, dP, dA] (dP / dA: canonical matrix entries in CSC order)."""
        self.init()
        P = params if isinstance(params, np.ndarray) else self.pack_params(params)
        P = np.ascontiguousarray(P, dtype=np.float64)
        sol_x = np.ascontiguousarray(sol_x, dtype=np.float64); sol_y = np.ascontiguousarray(sol_y, dtype=np.float64)
        B = sol_y.shape[0]
        if P.shape[0] == 1 and B > 1:
            P = np.ascontiguousarray(np.broadcast_to(P, (B, P.shape[1])))
        D = np.ascontiguousarray(dprim if isinstance(dprim, np.ndarray) else self.pack_dprim(dprim, B), dtype=np.float64)
        d = self.dims
        self._expect(P, (B, d.n_param), 'params'); self._expect(sol_x, (B, d.n_var), 'sol_x')
        self._expect(sol_y, (B, d.n_con), 'sol_y'); self._expect(D, (B, d.n_prim), 'dprim')
        dpar = np.empty((B, d.n_param))
        dq = np.empty((B, d.n_var)) if return_canonical else None
        dl = np.empty((B, d.n_con)) if return_canonical else None
        du = np.empty((B, d.n_con)) if return_canonical else None
        dP = np.empty((B, self.meta['nnzP'])) if return_canonical else None
        dA = np.empty((B, self.meta['nnzA'])) if return_canonical else None
        p = lambda a: None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))
        self._check(self._fn('cpg_gradient_batch_host_mat')(C.c_int(B), p(P), p(sol_x), p(sol_y), p(D), p(dpar), p(dq), p(dl),
                                                            p(du), p(dP), p(dA)))
        res = self.unpack_dparams(dpar)
        return (res, dq, dl, du, dP, dA) if return_canonical else res

    def gradient_batch_device_mat(self, params, sol_x, sol_y, dprim, dparams=None):
        import torch
        self.init()
        B = sol_y.shape[0]
        if dparams is None:
            dparams = torch.empty((B, self.dims.n_param), dtype=torch.float64, device=sol_y.device)
        ptr = lambda t: C.c_void_p(0 if t is None else t.data_ptr())
        stream = C.c_void_p(torch.cuda.current_stream(sol_y.device).cuda_stream)
        self._check(self._fn('cpg_gradient_batch_device_mat')(C.c_int(B), ptr(params), ptr(sol_x), ptr(sol_y), ptr(dprim),
                                                              ptr(dparams), None, None, None, None, None, stream))
        return dparams

    def gradient_batch(self, sol_y, dprim, sol_x=None, return_canonical=False):
        """Host arrays.  Returns (dict name -> (B, size) gradients of the batched parameters[, dq, dl, du])."""
        self.init()
        sol_y = np.ascontiguousarray(sol_y, dtype=np.float64)
        B = sol_y.shape[0]
        D = np.ascontiguousarray(dprim if isinstance(dprim, np.ndarray) else self.pack_dprim(dprim, B), dtype=np.float64)
        d = self.dims
        self._expect(sol_y, (B, d.n_con), 'sol_y'); self._expect(D, (B, d.n_prim), 'dprim')
        dpar = np.empty((B, d.n_param))
        dq = np.empty((B, d.n_var)) if return_canonical else None
        dl = np.empty((B, d.n_con)) if return_canonical else None
        du = np.empty((B, d.n_con)) if return_canonical else None
        p = lambda a: None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))
        self._check(self._fn('cpg_gradient_batch_host')(C.c_int(B), None, p(sol_y), p(D), p(dpar), p(dq), p(dl), p(du)))
        res = self.unpack_dparams(dpar)
        return (res, dq, dl, du) if return_canonical else res

    def gradient_batch_pinned(self, sol_y, dprim, dparams):
        """Host-buffer backward entry on caller-owned (ideally pinned) torch CPU tensors: H2D of sol_y / dprim, the backward
        kernel, D2H of dparams -- no allocation or copy on the Python side."""
        self.init()
        d = self.dims
        B = sol_y.shape[0]
        self._expect(sol_y, (B, d.n_con), 'sol_y'); self._expect(dprim, (B, d.n_prim), 'dprim'); self._expect(dparams, (B, d.n_param), 'dparams')
        ptr = lambda t: C.c_void_p(0 if t is None else t.data_ptr())
        self._check(self._fn('cpg_gradient_batch_host')(C.c_int(B), None, ptr(sol_y), ptr(dprim), ptr(dparams), None, None, None))
        return dparams

    def gradient_batch_device(self, sol_y, dprim, dparams=None):
        import torch
        self.init()
        B = sol_y.shape[0]
        if dparams is None:
            dparams = torch.empty((B, self.dims.n_param), dtype=torch.float64, device=sol_y.device)
        ptr = lambda t: C.c_void_p(0 if t is None else t.data_ptr())
        stream = C.c_void_p(torch.cuda.current_stream(sol_y.device).cuda_stream)
        self._check(self._fn('cpg_gradient_batch_device')(C.c_int(B), None, ptr(sol_y), ptr(dprim), ptr(dparams),
                                                          None, None, None, stream))
        return dparams

    # ---- reference-compatible single-instance entry: cpg_module.solve(upd, par)
    def solve(self, upd, par):
        vals = {}
        for p in self.meta['params']:
            if p['batched']:
                vals[p['name']] = np.atleast_1d(np.asarray(getattr(par, p['name']), dtype=np.float64)).reshape(1, -1)
            elif getattr(upd, p['name'], False):
                raise ValueError(f"parameter {p['name']} is shared in this generated code; regenerate or use update_shared_params")
        r = self.solve_batch(vals, return_canonical=True)
        self._last = dict(params=self.pack_params(vals), sol_x=r.sol_x, sol_y=r.sol_y)     # what gradient() differentiates at
        prim = SimpleNamespace(**{k: (v[0].flatten(order='F').tolist() if v[0].size > 1 else float(v[0].ravel()[0]))
                                  for k, v in r.cpg_prim.items()})
        dual = SimpleNamespace(**{k: (v[0].tolist() if v[0].size > 1 else float(v[0].ravel()[0])) for k, v in r.cpg_dual.items()})
        info = SimpleNamespace(obj_val=float(r.cpg_info.obj_val[0]), iter=int(r.cpg_info.iter[0]),
                               status=STATUS_STRINGS.get(int(r.cpg_info.status[0]), 'unknown'),
                               pri_res=float(r.cpg_info.pri_res[0]), dua_res=float(r.cpg_info.dua_res[0]),
                               time=r.cpg_info.time, gradient_primal=r.sol_x[0].tolist(), gradient_dual=r.sol_y[0].tolist())
        return SimpleNamespace(cpg_prim=prim, cpg_dual=dual, cpg_info=info)


# ----------------------------------------------------------------------------------------------------------------------
# IPM-CUDA libraries (SOCP families): include/cpg_b200_socp.h
ECOS_STATUS_STRINGS = {0: 'optimal', 1: 'primal infeasible', 2: 'dual infeasible', 10: 'optimal inaccurate',
                       11: 'primal infeasible inaccurate', 12: 'dual infeasible inaccurate', -1: 'maximum iterations reached',
                       -2: 'numerical problems (unreliable search direction)', -3: 'numerical problems (slacks or multipliers outside cone)',
                       -4: 'interrupted by signal or CTRL-C', -7: 'unknown problem in solver'}


class CpgB200SocpSettings(C.Structure):
    _fields_ = [('maxit', C.c_int), ('pad_', C.c_int), ('feastol', C.c_double), ('abstol', C.c_double),
                ('reltol', C.c_double), ('feastol_inacc', C.c_double), ('abstol_inacc', C.c_double),
                ('reltol_inacc', C.c_double)]


class CpgB200SocpDims(C.Structure):
    _fields_ = [(n, C.c_int) for n in ('n_var', 'n_eq', 'n_ineq', 'n_lp', 'n_soc', 'n_param', 'n_prim', 'n_dual',
                                       'threads_per_cta', 'smem_bytes')]


class SocpModule(Module):
    """One loaded IPM-CUDA solver library bound to one CUDA device."""

    def __init__(self, code_dir, device=0):
        self.code_dir = os.path.abspath(code_dir)
        with open(os.path.join(self.code_dir, 'cpg_meta.json')) as f:
            self.meta = json.load(f)
        self.prefix = self.meta['prefix']
        path = os.path.join(self.code_dir, 'libcpg_b200.so')
        if not os.path.exists(path):
            raise RuntimeError(f'{path} is missing: run cvxpygen_b200.codegen_ipm.compile_ipm_code (nvcc, sm_100a) first; '
                               'there is no CPU fallback')
        self.lib = C.CDLL(path)
        self._fn('cpg_b200_last_error').restype = C.c_char_p
        self.dims = CpgB200SocpDims()
        self._check(self._fn('cpg_socp_dims')(C.byref(self.dims)))
        self.settings = CpgB200SocpSettings()
        self.set_solver_default_settings()
        self.device = device
        self._initialised = False

    def set_solver_default_settings(self):
        self._fn('cpg_socp_default_settings')(C.byref(self.settings))

    def set_solver_setting(self, name, value):
        name = {'max_iters': 'maxit'}.get(name, name)
        if name == 'pad_' or name not in dict(CpgB200SocpSettings._fields_):
            raise AttributeError(f'Solver setting "{name}" not available.')
        setattr(self.settings, name, value)

    def update_shared_params(self, values):
        """Change user parameters shared by the whole batch: the offline setup (equilibration, KKT base image, affine maps)
        is re-run on the host and both constant images are re-uploaded -- what ECOS_updateData does for the reference
        (cvxpygen/solvers/ecos.py:88-117).  The compiled schedule must stay valid: same sparsity structure, same constant
        objective offset.  Needs the cvxpygen_b200 package."""
        import pickle
        from cvxpygen_b200 import codegen_ipm
        from cvxpygen_b200.offline.socp_setup import setup_socp_family
        with open(os.path.join(self.code_dir, 'cpg_family.pkl'), 'rb') as f:
            saved = pickle.load(f)
        fam = saved['family']
        if not hasattr(self, '_theta'):
            self._theta = fam.theta_default()
        for name, val in values.items():
            p = fam.param(name)
            if name in saved['batch_params']:
                raise ValueError(f'{name} is a batched parameter: pass it per instance to solve_batch')
            v = np.asarray(val, dtype=float)
            v = v.flatten(order='F') if v.ndim > 1 else v.ravel()
            if v.size != p.size:
                raise ValueError(f'parameter {name} stores {p.size} entries, got {v.size}')
            self._theta[p.col:p.col + p.size] = v
        st = setup_socp_family(fam, saved['batch_params'], theta=self._theta, threads=saved['threads'])
        strip = lambda txt: '\n'.join(txt.split('\n')[1:])            # first line carries the generation date
        with open(os.path.join(self.code_dir, 'c', 'include', 'cpg_ipm_family.h')) as f:
            compiled = f.read()
        if strip(codegen_ipm.family_header(st, self.prefix)) != strip(compiled):
            raise RuntimeError('the new parameter values change the compiled schedule (sparsity structure or objective offset): '
                               'regenerate the code (cpg.generate_code)')
        self.init()
        b = lambda x: (C.c_char_p(x), C.c_int(len(x)))
        self._check(self._fn('cpg_socp_load_constants')(*b(st.smem_blob), *b(st.gmem_blob)))
        return st

    def _qp_only(self, *a, **k):
        raise NotImplementedError('IPM-CUDA libraries carry no QP backward pass / warm start: the conic gradient goes through '
                                  'cvxpygen_b200.conic_grad (two-stage route, cvxpygen/canonicalizer.py:54-65)')
    gradient_batch = gradient_batch_device = gradient_batch_mat = gradient_batch_device_mat = _qp_only

    def solve_batch(self, params, return_canonical=False, **settings):
        self.init()
        for k, v in settings.items():
            self.set_solver_setting(k, v)
        P = params if isinstance(params, np.ndarray) else self.pack_params(params)
        P = np.ascontiguousarray(P, dtype=np.float64)
        d = self.dims
        self._expect(P, (None, d.n_param), 'params')
        B = P.shape[0]
        prim = np.empty((B, d.n_prim)); dual = np.empty((B, d.n_dual))
        x = np.empty((B, d.n_var)) if return_canonical else None
        y = np.empty((B, d.n_eq)) if return_canonical else None
        z = np.empty((B, d.n_ineq)) if return_canonical else None
        s = np.empty((B, d.n_ineq)) if return_canonical else None
        obj = np.empty(B); pri = np.empty(B); dua = np.empty(B)
        it = np.empty(B, dtype=np.int32); st = np.empty(B, dtype=np.int32)

        def p(a, t=C.c_double):
            return None if a is None else a.ctypes.data_as(C.POINTER(t))
        t0 = time.perf_counter()
        self._check(self._fn('cpg_socp_solve_batch_host')(C.c_int(B), p(P), p(prim), p(dual), p(x), p(y), p(z), p(s), p(obj),
                                                          p(it, C.c_int), p(st, C.c_int), p(pri), p(dua), C.byref(self.settings)))
        t1 = time.perf_counter()
        pr, du = self.unpack(prim, dual)
        info = SimpleNamespace(obj_val=obj, iter=it, status=st, pri_res=pri, dua_res=dua, time=t1 - t0)
        return SimpleNamespace(cpg_prim=pr, cpg_dual=du, cpg_info=info, prim=prim, dual=dual, sol_x=x, sol_y=y, sol_z=z, sol_s=s)

    def solve_batch_multi(self, params, devices=None, return_canonical=False, **settings):
        """The batch on several devices of this node from one call (cpg_socp_solve_batch_host_multi): contiguous shards, one
        host thread per device inside the library.  Same result object as solve_batch."""
        for k, v in settings.items():
            self.set_solver_setting(k, v)
        if devices is None:
            import torch
            devices = list(range(torch.cuda.device_count()))
        if not devices:
            raise RuntimeError('no CUDA device visible (there is no CPU fallback)')
        P = params if isinstance(params, np.ndarray) else self.pack_params(params)
        P = np.ascontiguousarray(P, dtype=np.float64)
        d = self.dims
        self._expect(P, (None, d.n_param), 'params')
        B = P.shape[0]
        prim = np.empty((B, d.n_prim)); dual = np.empty((B, d.n_dual))
        x = np.empty((B, d.n_var)) if return_canonical else None
        y = np.empty((B, d.n_eq)) if return_canonical else None
        z = np.empty((B, d.n_ineq)) if return_canonical else None
        s = np.empty((B, d.n_ineq)) if return_canonical else None
        obj = np.empty(B); pri = np.empty(B); dua = np.empty(B)
        it = np.empty(B, dtype=np.int32); st = np.empty(B, dtype=np.int32)
        dev = (C.c_int * len(devices))(*devices)

        def p(a, t=C.c_double):
            return None if a is None else a.ctypes.data_as(C.POINTER(t))
        t0 = time.perf_counter()
        self._check(self._fn('cpg_socp_solve_batch_host_multi')(C.c_int(len(devices)), dev, C.c_int(B), p(P), p(prim), p(dual), p(x), p(y),
                                                                p(z), p(s), p(obj), p(it, C.c_int), p(st, C.c_int), p(pri), p(dua),
                                                                C.byref(self.settings)))
        t1 = time.perf_counter()
        pr, du = self.unpack(prim, dual)
        info = SimpleNamespace(obj_val=obj, iter=it, status=st, pri_res=pri, dua_res=dua, time=t1 - t0)
        return SimpleNamespace(cpg_prim=pr, cpg_dual=du, cpg_info=info, prim=prim, dual=dual, sol_x=x, sol_y=y, sol_z=z, sol_s=s)

    def solve_batch_pinned(self, params, out):
        """Host-buffer entry on caller-owned (ideally pinned) torch CPU tensors, no allocation on the Python side.
        out: dict with prim, dual, obj, pri, dua (float64) and it, st (int32); optional sol_x, sol_y, sol_z, sol_s."""
        self.init()
        ptr = lambda t: C.c_void_p(0 if t is None else t.data_ptr())
        self._check(self._fn('cpg_socp_solve_batch_host')(C.c_int(params.shape[0]), ptr(params), ptr(out['prim']), ptr(out['dual']),
                                                          ptr(out.get('sol_x')), ptr(out.get('sol_y')), ptr(out.get('sol_z')),
                                                          ptr(out.get('sol_s')), ptr(out['obj']), ptr(out['it']), ptr(out['st']),
                                                          ptr(out['pri']), ptr(out['dua']), C.byref(self.settings)))
        return out

    def solve_batch_device(self, params, out=None, return_canonical=False, **_):
        import torch
        self.init()
        assert params.is_cuda and params.dtype == torch.float64 and params.is_contiguous()
        B = params.shape[0]
        d = self.dims
        dev = params.device
        if out is None:
            mk = lambda *shape, dtype=torch.float64: torch.empty(shape, dtype=dtype, device=dev)
            out = SimpleNamespace(prim=mk(B, d.n_prim), dual=mk(B, d.n_dual),
                                  sol_x=mk(B, d.n_var) if return_canonical else None, sol_y=mk(B, d.n_eq) if return_canonical else None,
                                  sol_z=mk(B, d.n_ineq) if return_canonical else None, sol_s=mk(B, d.n_ineq) if return_canonical else None,
                                  obj_val=mk(B), pri_res=mk(B), dua_res=mk(B), iter=mk(B, dtype=torch.int32), status=mk(B, dtype=torch.int32))
        ptr = lambda t: C.c_void_p(0 if t is None else t.data_ptr())
        stream = torch.cuda.current_stream(dev).cuda_stream
        self._check(self._fn('cpg_socp_solve_batch_device')(C.c_int(B), ptr(params), ptr(out.prim), ptr(out.dual), ptr(out.sol_x),
                                                            ptr(out.sol_y), ptr(out.sol_z), ptr(out.sol_s), ptr(out.obj_val),
                                                            ptr(out.iter), ptr(out.status), ptr(out.pri_res), ptr(out.dua_res),
                                                            C.byref(self.settings), C.c_void_p(stream)))
        return out

    def solve(self, upd, par):
        """b3: cpg_module.solve(upd, par) -> result, a batch of one (status is ECOS's integer exit flag, like the
        reference's ECOS path: status_is_int, cvxpygen/solvers/ecos.py:29)."""
        vals = {p['name']: np.asarray(getattr(par, p['name']), dtype=float) for p in self.meta['params'] if p['batched']}
        r = self.solve_batch(vals)
        prim = SimpleNamespace(**{k: (v[0].flatten(order='F').tolist() if v[0].size > 1 else float(v[0].ravel()[0]))
                                  for k, v in r.cpg_prim.items()})
        dual = SimpleNamespace(**{k: (v[0].tolist() if v[0].size > 1 else float(v[0].ravel()[0])) for k, v in r.cpg_dual.items()})
        info = SimpleNamespace(obj_val=float(r.cpg_info.obj_val[0]), iter=int(r.cpg_info.iter[0]), status=int(r.cpg_info.status[0]),
                               pri_res=float(r.cpg_info.pri_res[0]), dua_res=float(r.cpg_info.dua_res[0]), time=r.cpg_info.time)
        return SimpleNamespace(cpg_prim=prim, cpg_dual=dual, cpg_info=info)


_modules = {}


def load(code_dir=None, device=0) -> Module:
    code_dir = os.path.dirname(os.path.abspath(__file__)) if code_dir is None else os.path.abspath(code_dir)
    key = (code_dir, device)
    if key not in _modules:
        with open(os.path.join(code_dir, 'cpg_meta.json')) as f:
            solver = json.load(f).get('solver', 'ADMM-CUDA')
        _modules[key] = (SocpModule if solver == 'IPM-CUDA' else Module)(code_dir, device)
    return _modules[key]
