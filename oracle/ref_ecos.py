"""oracle/ref_ecos.py -- TEST INFRASTRUCTURE (oracle), not product code.
ctypes binding of oracle/_ref/libecos_ref.so: the unmodified vendored ECOS 2.0.8 (reference:
cvxpygen/solvers/ecos/src/ecos.c:1075 ECOS_solve, :1648 ECOS_updateData) behind oracle/ecos_ref_driver.c."""
import ctypes as C
import os

import numpy as np
import scipy.sparse as sp

_LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref', 'libecos_ref.so')
EXIT = {0: 'optimal', 1: 'primal infeasible', 2: 'dual infeasible', 10: 'optimal inaccurate', -1: 'maxit', -2: 'numerics'}


def available():
    return os.path.exists(_LIB)


class RefECOS:
    def __init__(self, c, A, b, G, h, l, q, feastol=1e-8, abstol=1e-8, reltol=1e-8, maxit=100):
        self.lib = C.CDLL(_LIB)
        self.lib.ecos_ref_setup.restype = C.c_void_p
        self.lib.ecos_ref_solve_batch.restype = C.c_double
        A = sp.csc_matrix(A); G = sp.csc_matrix(G); A.sort_indices(); G.sort_indices()
        self.n, self.p, self.m = G.shape[1], A.shape[0], G.shape[0]
        L = lambda a: np.ascontiguousarray(a, dtype=np.int64)
        D = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        self._keep = [L(q), D(G.data), L(G.indptr), L(G.indices), D(A.data), L(A.indptr), L(A.indices), D(c), D(h), D(b)]
        pl = lambda a: a.ctypes.data_as(C.POINTER(C.c_long))
        pd = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        k = self._keep
        self.h_ = C.c_void_p(self.lib.ecos_ref_setup(C.c_long(self.n), C.c_long(self.m), C.c_long(self.p), C.c_long(l),
                                                      C.c_long(len(q)), pl(k[0]), pd(k[1]), pl(k[2]), pl(k[3]), pd(k[4]), pl(k[5]),
                                                      pl(k[6]), pd(k[7]), pd(k[8]), pd(k[9]), C.c_double(feastol),
                                                      C.c_double(abstol), C.c_double(reltol), C.c_long(maxit)))
        if not self.h_:
            raise RuntimeError('ECOS_setup failed')

    def solve_batch(self, c=None, h=None, b=None, B=None, G=None, A=None):
        """c (B, n), h (B, m), b (B, p): per-instance vectors; G (B, nnz(G)), A (B, nnz(A)): per-instance matrix VALUES in the CSC order
        of the matrices given to the constructor (the reference re-equilibrates on every ECOS_updateData)."""
        for a in (c, h, b, G, A):
            if a is not None:
                B = np.asarray(a).shape[0]
        B = B or 1
        D = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)
        c, h, b, G, A = D(c), D(h), D(b), D(G), D(A)
        if G is not None or A is not None:
            assert G is None or G.shape == (B, len(self._keep[1])), 'G values: one row of nnz(G) per instance'
            assert A is None or A.shape == (B, len(self._keep[4])), 'A values: one row of nnz(A) per instance'
            self.lib.ecos_ref_solve_batch_mat.restype = C.c_double
            x = np.zeros((B, self.n)); y = np.zeros((B, max(self.p, 1))); z = np.zeros((B, self.m)); s = np.zeros((B, self.m))
            pc = np.zeros(B); pr = np.zeros(B); dr = np.zeros(B); it = np.zeros(B, np.int64); ef = np.zeros(B, np.int64)
            pd = lambda a: None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))
            pl = lambda a: a.ctypes.data_as(C.POINTER(C.c_long))
            sec = self.lib.ecos_ref_solve_batch_mat(self.h_, C.c_long(B), pd(c), pd(h), pd(b), pd(G), pd(A), pd(x), pd(y), pd(z), pd(s),
                                                    pd(pc), pl(it), pl(ef), pd(pr), pd(dr))
            return dict(x=x, y=y[:, :self.p], z=z, s=s, pcost=pc, iter=it, exitflag=ef, pres=pr, dres=dr, seconds=sec)
        x = np.zeros((B, self.n)); y = np.zeros((B, max(self.p, 1))); z = np.zeros((B, self.m)); s = np.zeros((B, self.m))
        pc = np.zeros(B); pr = np.zeros(B); dr = np.zeros(B); it = np.zeros(B, np.int64); ef = np.zeros(B, np.int64)
        pd = lambda a: None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))
        pl = lambda a: a.ctypes.data_as(C.POINTER(C.c_long))
        sec = self.lib.ecos_ref_solve_batch(self.h_, C.c_long(B), pd(c), pd(h), pd(b), pd(x), pd(y), pd(z), pd(s),
                                            pd(pc), pl(it), pl(ef), pd(pr), pd(dr))
        return dict(x=x, y=y[:, :self.p], z=z, s=s, pcost=pc, iter=it, exitflag=ef, pres=pr, dres=dr, seconds=sec)
