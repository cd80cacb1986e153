"""oracle/ref_osqp.py -- TEST INFRASTRUCTURE (oracle), not product code.

ctypes binding of oracle/_ref/libosqp_ref.so: the unmodified vendored OSQP 0.6.2
(reference: cvxpygen/solvers/osqp-python/osqp_sources/src/osqp.c:76 osqp_setup,
:288 osqp_solve, :752 osqp_update_lin_cost, :784 osqp_update_bounds) behind the
batch driver oracle/osqp_ref_driver.c.
"""
import ctypes as C
import os
import numpy as np
import scipy.sparse as sp

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, '_ref', 'libosqp_ref.so')

STATUS = {1: 'solved', 2: 'solved inaccurate', 3: 'primal infeasible inaccurate',
          4: 'dual infeasible inaccurate', -2: 'maximum iterations reached',
          -3: 'primal infeasible', -4: 'dual infeasible', -7: 'problem non convex',
          -10: 'unsolved'}  # osqp_sources/include/constants.h:18-30


class RefSettings(C.Structure):
    _fields_ = [('max_iter', C.c_int),
                ('eps_abs', C.c_double), ('eps_rel', C.c_double),
                ('eps_prim_inf', C.c_double), ('eps_dual_inf', C.c_double),
                ('rho', C.c_double), ('sigma', C.c_double), ('alpha', C.c_double),
                ('scaling', C.c_int), ('adaptive_rho', C.c_int), ('adaptive_rho_interval', C.c_int),
                ('adaptive_rho_tolerance', C.c_double),
                ('scaled_termination', C.c_int), ('check_termination', C.c_int),
                ('warm_start', C.c_int), ('polish', C.c_int)]


def available():
    return os.path.exists(_LIB_PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(_LIB_PATH)
        L.ref_osqp_setup.restype = C.c_void_p
        L.ref_osqp_solve_batch.restype = C.c_double
        L.ref_osqp_solve_batch_mat.restype = C.c_double
        L.ref_osqp_adaptive_rho_interval.restype = C.c_int
        _lib = L
    return _lib


def _p(a, t):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


class RefOSQP:
    """One OSQP 0.6.2 workspace per host thread, set up once per family."""

    def __init__(self, P, q, A, l, u, nthreads=1, **settings):
        L = lib()
        P = sp.triu(sp.csc_matrix(P), format='csc'); P.sort_indices()
        A = sp.csc_matrix(A); A.sort_indices()
        self.n, self.m = P.shape[0], A.shape[0]
        s = RefSettings()
        L.ref_osqp_default_settings(C.byref(s))
        for k, v in settings.items():
            if not hasattr(s, k):
                raise AttributeError(f'unknown OSQP setting {k}')
            setattr(s, k, v)
        self.settings = s
        q = np.ascontiguousarray(q, dtype=np.float64)
        l = np.ascontiguousarray(np.clip(l, -1e30, 1e30), dtype=np.float64)
        u = np.ascontiguousarray(np.clip(u, -1e30, 1e30), dtype=np.float64)
        self._keep = (P, A, q, l, u)
        Pp, Pi, Px = P.indptr.astype(np.int32), P.indices.astype(np.int32), P.data.astype(np.float64)
        Ap, Ai, Ax = A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.astype(np.float64)
        self.nthreads = max(1, int(nthreads))
        self.h = L.ref_osqp_setup(C.c_int(self.n), C.c_int(self.m),
                                  _p(Pp, C.c_int), _p(Pi, C.c_int), _p(Px, C.c_double), _p(q, C.c_double),
                                  _p(Ap, C.c_int), _p(Ai, C.c_int), _p(Ax, C.c_double),
                                  _p(l, C.c_double), _p(u, C.c_double), C.byref(s), C.c_int(self.nthreads))
        if not self.h:
            raise RuntimeError('osqp_setup failed')
        self.h = C.c_void_p(self.h)

    def scaling(self):
        D = np.zeros(self.n); E = np.zeros(self.m); c = C.c_double()
        lib().ref_osqp_get_scaling(self.h, _p(D, C.c_double), _p(E, C.c_double), C.byref(c))
        return D, E, c.value

    def adaptive_rho_interval(self):
        return lib().ref_osqp_adaptive_rho_interval(self.h)

    def solve_batch(self, q=None, l=None, u=None, B=None, x0=None, y0=None, nthreads=None):
        """q:(B,n) l,u:(B,m) or None. Returns dict with x,y,obj,iter,status,pri_res,dua_res,rho_updates,seconds."""
        for a in (q, l, u):
            if a is not None:
                B = a.shape[0]
        if B is None:
            B = 1
        def prep(a, clipinf=False):
            if a is None:
                return None
            a = np.ascontiguousarray(a, dtype=np.float64)
            return np.clip(a, -1e30, 1e30) if clipinf else a
        q, l, u = prep(q), prep(l, True), prep(u, True)
        x0, y0 = prep(x0), prep(y0)
        x = np.zeros((B, self.n)); y = np.zeros((B, self.m))
        obj = np.zeros(B); it = np.zeros(B, np.int32); st = np.zeros(B, np.int32)
        pr = np.zeros(B); dr = np.zeros(B); ru = np.zeros(B, np.int32)
        nt = self.nthreads if nthreads is None else nthreads
        sec = lib().ref_osqp_solve_batch(self.h, C.c_int(B), _p(q, C.c_double), _p(l, C.c_double), _p(u, C.c_double),
                                         _p(x0, C.c_double), _p(y0, C.c_double),
                                         _p(x, C.c_double), _p(y, C.c_double), _p(obj, C.c_double),
                                         _p(it, C.c_int), _p(st, C.c_int), _p(pr, C.c_double), _p(dr, C.c_double),
                                         _p(ru, C.c_int), C.c_int(nt))
        return dict(x=x, y=y, obj=obj, iter=it, status=st, pri_res=pr, dua_res=dr, rho_updates=ru, seconds=sec)

    def solve_batch_mat(self, Px=None, Ax=None, q=None, l=None, u=None, nthreads=None):
        """Per-instance matrices: Px (B, nnz(P upper)) / Ax (B, nnzA) in CSC order (either may be None), then q/l/u as in
        solve_batch.  Per instance: vectors back to their setup values, osqp_update_P_A (osqp.c:1158-1264: unscale,
        overwrite, scale_data, refactor), osqp_update_lin_cost / _bounds, osqp_solve."""
        B = next(a.shape[0] for a in (Px, Ax, q, l, u) if a is not None)
        c64 = lambda a, clip=False: None if a is None else np.ascontiguousarray(np.clip(a, -1e30, 1e30) if clip else a, dtype=np.float64)
        Px, Ax, q, l, u = c64(Px), c64(Ax), c64(q), c64(l, True), c64(u, True)
        x = np.zeros((B, self.n)); y = np.zeros((B, self.m))
        obj = np.zeros(B); it = np.zeros(B, np.int32); st = np.zeros(B, np.int32)
        pr = np.zeros(B); dr = np.zeros(B); ru = np.zeros(B, np.int32)
        nt = self.nthreads if nthreads is None else nthreads
        sec = lib().ref_osqp_solve_batch_mat(self.h, C.c_int(B), _p(Px, C.c_double), C.c_int(0 if Px is None else Px.shape[1]),
                                             _p(Ax, C.c_double), C.c_int(0 if Ax is None else Ax.shape[1]),
                                             _p(q, C.c_double), _p(l, C.c_double), _p(u, C.c_double),
                                             _p(x, C.c_double), _p(y, C.c_double), _p(obj, C.c_double),
                                             _p(it, C.c_int), _p(st, C.c_int), _p(pr, C.c_double), _p(dr, C.c_double),
                                             _p(ru, C.c_int), C.c_int(nt))
        return dict(x=x, y=y, obj=obj, iter=it, status=st, pri_res=pr, dua_res=dr, rho_updates=ru, seconds=sec)

    def __del__(self):
        try:
            if self.h:
                lib().ref_osqp_free(self.h)
        except Exception:
            pass
