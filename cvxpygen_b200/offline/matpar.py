"""numpy emulation of the matrix-parameter kernel's table-driven steps (csrc/matpar_kernel.cuh: matpar_prepare and the
per-instance KKT assembly of solve_instance<Fam, 2>), reading the SAME packed blob the kernel reads.  Used by the CPU
tests to pin the tables of offline/blob.py:pack_matpar_blob against a direct computation (ruiz_equilibrate + assemble_kkt),
the way offline/refactor.py:emulate_factor pins the refactorisation tables.
"""
import struct

import numpy as np

from .blob import MAT_HEADER_FIELDS, LANES

MIN_SCALING, MAX_SCALING = 1e-4, 1e4


class MatBlob:
    def __init__(self, blob: bytes):
        n_f = len(MAT_HEADER_FIELDS)
        vals = struct.unpack('<' + 'i' * n_f, blob[:4 * n_f])
        self.h = dict(zip([nm for _, nm in MAT_HEADER_FIELDS], vals))
        h = self.h
        self.I32 = np.frombuffer(blob, dtype='<i4', offset=h['off_i32'], count=(h['off_f64'] - h['off_i32']) // 4)
        self.F64 = np.frombuffer(blob, dtype='<f8', offset=h['off_f64'], count=(h['off_u16'] - h['off_f64']) // 8)
        self.U16 = np.frombuffer(blob, dtype='<u2', offset=h['off_u16'], count=(h['total_bytes'] - h['off_u16']) // 2)

    def f64(self, key, count):
        return self.F64[self.h[key]:self.h[key] + count]

    def u16(self, key, count):
        return self.U16[self.h[key]:self.h[key] + count].astype(np.int64)

    def ell_map(self, key, nnz, base_key, th):
        """base + map * theta over the entries (cpg_canonicalize_<P|A>)."""
        out = np.array(self.f64(base_key, nnz), dtype=float)
        for blk, e0 in enumerate(range(0, nnz, LANES)):
            K, fo, uo = self.I32[self.h[key] + 3 * blk: self.h[key] + 3 * blk + 3]
            for lane in range(min(LANES, nnz - e0)):
                acc = out[e0 + lane]
                for kk in range(K):
                    acc = acc + self.F64[fo + kk * LANES + lane] * th[self.U16[uo + kk * LANES + lane]]
                out[e0 + lane] = acc
        return out

    def ix_rows(self, key, n_rows):
        """list over rows of (entry numbers, operand positions) incl. padding entries."""
        rows = []
        for r in range(n_rows):
            blk, lane = divmod(r, LANES)
            K, io, co = self.I32[self.h[key] + 3 * blk: self.h[key] + 3 * blk + 3]
            rows.append((self.U16[io + lane + LANES * np.arange(K)].astype(np.int64),
                         self.U16[co + lane + LANES * np.arange(K)].astype(np.int64)))
        return rows


def _limit(v):
    v = np.where(v < MIN_SCALING, 1.0, v)
    return np.where(v > MAX_SCALING, MAX_SCALING, v)


def emulate_prepare(blob: bytes, th: np.ndarray):
    """Returns dict(Pv, Av, D, E, c) exactly as matpar_prepare leaves them (same operation order)."""
    mb = MatBlob(blob)
    h = mb.h
    n, m, nnzP, nnzA = h['n'], h['m'], h['nnzP'], h['nnzA']
    Pv = np.append(mb.ell_map('i_ellMP', nnzP, 'f_Pbase', th), 0.0)
    Av = np.append(mb.ell_map('i_ellMA', nnzA, 'f_Abase', th), 0.0)
    ixP, ixA, ixAt = mb.ix_rows('i_ixP', n), mb.ix_rows('i_ixA', m), mb.ix_rows('i_ixAt', n)
    Prow, Pcol = mb.u16('h_Prow', nnzP), mb.u16('h_Pcol', nnzP)
    Arow, Acol = mb.u16('h_Arow', nnzA), mb.u16('h_Acol', nnzA)
    absmax = lambda rows, vals: np.array([np.abs(vals[ix]).max() if len(ix) else 0.0 for ix, _ in rows])
    q = np.array(mb.f64('f_q_un', n), dtype=float)
    D, E, c = np.ones(n), np.ones(m), 1.0
    for _ in range(h['scaling_iters']):
        nr = absmax(ixP, Pv)
        if m:
            nr = np.maximum(nr, absmax(ixAt, Av))
        dt = 1.0 / np.sqrt(_limit(nr))
        et = 1.0 / np.sqrt(_limit(absmax(ixA, Av))) if m else np.zeros(0)
        Pv[:nnzP] = (Pv[:nnzP] * dt[Prow]) * dt[Pcol]
        Av[:nnzA] = (Av[:nnzA] * et[Arow]) * dt[Acol]
        q = q * dt; D = D * dt; E = E * et
        cn = absmax(ixP, Pv)
        mean = 0.0
        for t in cn.tolist():
            mean += t
        mean /= n
        nq = float(_limit(np.array([np.abs(q).max() if n else 0.0]))[0])
        ct = 1.0 / float(_limit(np.array([max(mean, nq)]))[0])
        Pv[:nnzP] *= ct; q = q * ct; c *= ct
    return dict(Pv=Pv[:nnzP], Av=Av[:nnzA], D=D, E=E, c=c, blob=mb)


def emulate_assemble(mb: MatBlob, Pv, Av, rho_slot, rho_vec):
    """S (slot order) as the MODE-2 refactor() builds it before tail_factor."""
    h = mb.h
    S = np.array(mb.f64('f_S0', h['n_slots']), dtype=float)
    np.add.at(S, mb.u16('h_Pslot', h['nnzP']), Pv)
    S[mb.u16('h_Aslot', h['nnzA'])] = Av
    S[np.asarray(rho_slot)] = -1.0 / np.asarray(rho_vec)
    return S
