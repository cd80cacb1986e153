"""Constants blob of one ADMM-CUDA problem family.

Everything a CTA needs besides the per-instance parameters is packed offline into ONE
contiguous, 16-byte aligned byte string which the kernel stages global -> shared with
TMA bulk copies (cp.async.bulk) once per CTA.  This is the role `workspace.c` plays for
the reference's embedded OSQP (data, scaling, L, Dinv, P, rho_inv_vec, ... --
osqp-python/module/codegen/utils.py:66-254 as listed in SURVEY Appendix A), re-laid-out
for a warp: lane-interleaved ELL tiles instead of CSC columns.

Layout:  [ header | int32 area | float64 area | uint16 area ]
The header field list below is the single source of truth; `header_struct_c()` emits the
matching C struct into the generated code, so host packer and kernel cannot drift apart.
"""
import struct
from typing import Dict, List, Tuple

import numpy as np
import scipy.sparse as sp

import threading

_TLS = threading.local()

MAGIC = 0x42323030   # 'B200'
LANES = 32

# (ctype, name) -- ints first, then doubles; total kept a multiple of 16 bytes
HEADER_FIELDS: List[Tuple[str, str]] = [
    ('int', 'magic'), ('int', 'total_bytes'), ('int', 'n'), ('int', 'm'),
    ('int', 'nk'), ('int', 'npb'), ('int', 'n_prim'), ('int', 'n_dual'),
    ('int', 'n_tiles'), ('int', 'n_fwd_tiles'), ('int', 'n_trail_tiles'), ('int', 'is_max'),
    ('int', 'off_i32'), ('int', 'off_f64'), ('int', 'off_u16'), ('int', 'n_theta_shared'),
    # int32-area indices
    ('int', 'i_tiles'), ('int', 'i_ellA'), ('int', 'i_ellAt'), ('int', 'i_ellP'),
    ('int', 'i_ellMq'), ('int', 'i_ellMl'), ('int', 'i_ellMu'), ('int', 'pad_i0'),
    # float64-area indices (in doubles)
    ('int', 'f_D'), ('int', 'f_Dinv'), ('int', 'f_E'), ('int', 'f_Einv'),
    ('int', 'f_qbase'), ('int', 'f_lbase'), ('int', 'f_ubase'), ('int', 'pad_f0'),
    # uint16-area indices
    ('int', 'h_pinvx'), ('int', 'h_pinvz'), ('int', 'h_ctype'), ('int', 'h_prim'),
    ('int', 'h_dual'), ('int', 'n_bq'), ('int', 'n_bc'), ('int', 'nb_slots'),
    # two-instances-per-warp kernel: encoded tiles, scaled base vectors, accessor tables, batched-row maps
    ('int', 'i_tiles2'), ('int', 'f_qs'), ('int', 'f_ls'), ('int', 'f_us'),
    ('int', 'i_addr_q'), ('int', 'i_addr_l'), ('int', 'i_addr_u'), ('int', 'h_bq_row'),
    ('int', 'h_bc_row'), ('int', 'i_bq_ptr'), ('int', 'i_bl_ptr'), ('int', 'i_bu_ptr'),
    ('int', 'h_bq_col'), ('int', 'h_bl_col'), ('int', 'h_bu_col'), ('int', 'f_bq_val'),
    ('int', 'f_bl_val'), ('int', 'f_bu_val'), ('int', 'pad_v0'), ('int', 'pad_v1'),
    ('double', 'sigma'), ('double', 'rho'), ('double', 'c'), ('double', 'cinv'),
    ('double', 'd_const'), ('double', 'reserved0'),
]


def header_struct_c(name='CpgBlobHeader') -> str:
    lines = [f'struct {name} {{']
    for t, n in HEADER_FIELDS:
        lines.append(f'  {t} {n};')
    lines.append('};')
    return '\n'.join(lines) + '\n'


def _pack_header(vals: Dict[str, float]) -> bytes:
    fmt = '<' + ''.join('i' if t == 'int' else 'd' for t, _ in HEADER_FIELDS)
    data = struct.pack(fmt, *[(int(vals.get(n, 0)) if t == 'int' else float(vals.get(n, 0.0))) for t, n in HEADER_FIELDS])
    assert len(data) % 16 == 0, len(data)
    return data


class _Areas:
    def __init__(self):
        self.i32: List[int] = []
        self.f64: List[float] = []
        self.u16: List[int] = []

    def add_i32(self, arr) -> int:
        off = len(self.i32); self.i32.extend(int(v) for v in np.asarray(arr).ravel()); return off

    def add_f64(self, arr) -> int:
        off = len(self.f64); self.f64.extend(float(v) for v in np.asarray(arr, dtype=float).ravel()); return off

    def add_u16(self, arr) -> int:
        a = np.asarray(arr).ravel()
        assert a.size == 0 or (a.min() >= 0 and a.max() < 65536)
        off = len(self.u16); self.u16.extend(int(v) for v in a); return off


def ell_row_blocks(Mcsr: sp.csr_matrix, n_rows: int):
    """Row-blocked ELL of a CSR matrix: block rb holds rows 32*rb .. 32*rb+31 (row = lane + 32*rb),
    K_rb = longest row in the block; entry k of a row sits at [k*32 + lane].
    Returns list of (K, vals(K,32), cols(K,32))."""
    Mcsr = sp.csr_matrix(Mcsr); Mcsr.sort_indices()
    blocks = []
    for r0 in range(0, n_rows, LANES):
        rows = range(r0, min(r0 + LANES, n_rows))
        lens = [Mcsr.indptr[r + 1] - Mcsr.indptr[r] for r in rows]
        K = max(lens) if lens else 0
        vals = np.zeros((K, LANES)); cols = np.zeros((K, LANES), dtype=np.int64)
        for lane, r in enumerate(rows):
            s, e = Mcsr.indptr[r], Mcsr.indptr[r + 1]
            vals[:e - s, lane] = Mcsr.data[s:e]
            cols[:e - s, lane] = Mcsr.indices[s:e]
        blocks.append((K, vals, cols))
    return blocks


def _add_ell(ar: _Areas, blocks) -> int:
    """int32 table: per block [K, f64 offset, u16 offset]; returns table index."""
    table = []
    for K, vals, cols in blocks:
        table += [K, ar.add_f64(vals), ar.add_u16(cols)]
    return ar.add_i32(table) if table else ar.add_i32([0, 0, 0])


def last_solve_source() -> str:
    """generated straight-line KKT solve of the last pack_blob(with_tiles=True) call of THIS thread"""
    return _TLS.last_solve_source


def pack_blob(*, n, m, perm, schedule, Ps_upper, As, D, E, c, sigma, rho, ctype,
              q_base, l_base, u_base, Mq_b, Ml_b, Mu_b, npb, prim_idx, dual_idx,
              d_const, is_max, with_tiles=True) -> bytes:
    """All vectors are in natural (unpermuted) canonical order.  q/l/u_base are the UNSCALED
    canonical vectors with the batched parameters set to zero (shared parameters and constants
    folded in); M*_b are the CSR maps restricted to the batched-parameter columns."""
    nk = n + m
    ar = _Areas()
    pinv = np.empty(nk, dtype=np.int64); pinv[np.asarray(perm)] = np.arange(nk)
    # --- KKT-solve tiles, encoded for the pair kernel (sparse gather / dense broadcast; offline/schedule.py)
    from .schedule import encode_best
    tile_tab = []
    tile_info, encodings = [], []
    for t in (schedule.tiles if with_tiles else []):
        e = encode_best(t, nk)
        encodings.append(e)
        f = ar.add_f64(e['vals'])
        hr = ar.add_u16(np.concatenate([t.rows, np.zeros(LANES - len(t.rows), dtype=np.uint16)]))
        if e['kind'] == 0:
            # packed words hold BYTE offsets into the interleaved pair vector (pos * 16 < 65536)
            words = np.asarray(e['cols32'], dtype=np.uint32)
            words = ((words & 0xffff) * 16) | (((words >> 16) * 16) << 16)
            assert nk * 16 < 65536
            aux = ar.add_i32(words.astype(np.uint32).view(np.int32))
            tile_tab += [0, f, aux, e['K'], t.r_pad, len(t.rows), hr, 0]
        else:
            aux = ar.add_i32(np.asarray(e['segs'], dtype=np.int64).ravel())
            tile_tab += [1, f, aux, e['K'], t.r_pad, len(t.rows), hr, len(e['segs'])]
        tile_info.append(dict(f64=f, aux=aux, rows=hr))
    while len(ar.i32) % 4:
        ar.i32.append(0)
    i_tiles2 = ar.add_i32(tile_tab if tile_tab else [0] * 8)
    i_tiles = i_tiles2
    # --- residual operators (scaled data): A (rows), A' (rows = variables), full symmetric P
    As = sp.csr_matrix(As)
    Pfull = sp.csr_matrix(Ps_upper + sp.triu(Ps_upper, 1).T)
    i_A = _add_ell(ar, ell_row_blocks(As, m))
    # A' gathers y from w[n + j]
    At = sp.csr_matrix(As.T)
    At_blocks = [(K, v, cidx + n) for K, v, cidx in ell_row_blocks(At, n)]
    i_At = _add_ell(ar, At_blocks)
    i_P = _add_ell(ar, ell_row_blocks(Pfull, n))
    i_Mq = _add_ell(ar, ell_row_blocks(Mq_b, n))
    i_Ml = _add_ell(ar, ell_row_blocks(Ml_b, m))
    i_Mu = _add_ell(ar, ell_row_blocks(Mu_b, m))
    D = np.asarray(D, dtype=float); E = np.asarray(E, dtype=float)
    Mq_c, Ml_c, Mu_c = sp.csr_matrix(Mq_b), sp.csr_matrix(Ml_b), sp.csr_matrix(Mu_b)
    bq_rows = np.nonzero(np.diff(Mq_c.indptr))[0] if npb else np.zeros(0, dtype=int)
    bc_rows = np.nonzero(np.diff(Ml_c.indptr) + np.diff(Mu_c.indptr))[0] if npb else np.zeros(0, dtype=int)
    n_bq, n_bc = len(bq_rows), len(bc_rows)
    f_qs = ar.add_f64((D * np.asarray(q_base)) * c)
    f_ls = ar.add_f64(E * np.clip(l_base, -1e30, 1e30))
    f_us = ar.add_f64(E * np.clip(u_base, -1e30, 1e30))

    def addr_table(f_base, nrows, brow, slot0, stride):
        tab = np.array([f_base + r for r in range(nrows)], dtype=np.int64)          # >= 0: index into the f64 area
        for k, r in enumerate(brow):
            tab[r] = -(slot0 + stride * k) - 1                                      # < 0: per-warp slot of a batched row
        return tab
    # per-warp slots: [q rows][l rows][u rows], each slot holds the value for the two instances of the warp
    addr_q = addr_table(f_qs, n, bq_rows, 0, 1)
    addr_l = addr_table(f_ls, m, bc_rows, n_bq, 1)
    addr_u = addr_table(f_us, m, bc_rows, n_bq + n_bc, 1)

    def csr_rows(M, rows):
        ptr, cols, vals = [0], [], []
        for r in rows:
            s_, e_ = M.indptr[r], M.indptr[r + 1]
            cols += M.indices[s_:e_].tolist(); vals += M.data[s_:e_].tolist(); ptr.append(len(cols))
        return ptr, cols, vals
    pq, cq, vq = csr_rows(Mq_c, bq_rows)
    pl, cl, vl = csr_rows(Ml_c, bc_rows)
    pu, cu, vu = csr_rows(Mu_c, bc_rows)
    v2 = dict(i_tiles2=i_tiles2, n_bq=n_bq, n_bc=n_bc, nb_slots=n_bq + 2 * n_bc, f_qs=f_qs, f_ls=f_ls, f_us=f_us,
              i_addr_q=ar.add_i32(addr_q), i_addr_l=ar.add_i32(addr_l), i_addr_u=ar.add_i32(addr_u),
              h_bq_row=ar.add_u16(bq_rows), h_bc_row=ar.add_u16(bc_rows),
              i_bq_ptr=ar.add_i32(pq), i_bl_ptr=ar.add_i32(pl), i_bu_ptr=ar.add_i32(pu),
              h_bq_col=ar.add_u16(cq), h_bl_col=ar.add_u16(cl), h_bu_col=ar.add_u16(cu),
              f_bq_val=ar.add_f64(vq), f_bl_val=ar.add_f64(vl), f_bu_val=ar.add_f64(vu))
    hv = dict(magic=MAGIC, n=n, m=m, nk=nk, npb=npb, n_prim=len(prim_idx), n_dual=len(dual_idx),
              n_tiles=len(schedule.tiles), n_fwd_tiles=schedule.n_fwd_tiles,
              n_trail_tiles=schedule.n_trailing_tiles, is_max=int(is_max),
              i_tiles=i_tiles, i_ellA=i_A, i_ellAt=i_At, i_ellP=i_P, i_ellMq=i_Mq, i_ellMl=i_Ml, i_ellMu=i_Mu,
              f_D=ar.add_f64(D), f_Dinv=ar.add_f64(1.0 / np.asarray(D)), f_E=ar.add_f64(E),
              f_Einv=ar.add_f64(1.0 / np.asarray(E)),
              f_qbase=ar.add_f64(q_base), f_lbase=ar.add_f64(l_base), f_ubase=ar.add_f64(u_base),
              h_pinvx=ar.add_u16(pinv[:n]), h_pinvz=ar.add_u16(pinv[n:]),
              h_ctype=ar.add_u16(np.asarray(ctype) + 1),      # stored as 0 loose / 1 ineq / 2 eq
              h_prim=ar.add_u16(prim_idx), h_dual=ar.add_u16(dual_idx),
              sigma=sigma, rho=rho, c=c, cinv=1.0 / c, d_const=d_const)
    hv.update(v2)
    hdr_len = len(_pack_header(hv))

    def pad16(b: bytes) -> bytes:
        return b + b'\0' * ((-len(b)) % 16)
    i32b = pad16(np.asarray(ar.i32, dtype='<i4').tobytes())
    f64b = pad16(np.asarray(ar.f64, dtype='<f8').tobytes())
    u16b = pad16(np.asarray(ar.u16, dtype='<u2').tobytes())
    hv['off_i32'] = hdr_len
    hv['off_f64'] = hdr_len + len(i32b)
    hv['off_u16'] = hv['off_f64'] + len(f64b)
    hv['total_bytes'] = hv['off_u16'] + len(u16b)
    blob = _pack_header(hv) + i32b + f64b + u16b
    assert len(blob) == hv['total_bytes'] and len(blob) % 16 == 0
    if with_tiles:
        from .emit_solve import emit_kkt_solve
        _TLS.last_solve_source = emit_kkt_solve(schedule, encodings, tile_info)   # per thread: families are built concurrently
        _TLS.last_header = dict(hv)
    return blob


# ---------------------------------------------------------------------------------------------------------
# Tail blob: tables of offline/refactor.py for the in-kernel numeric re-factorisation.  Lives in GLOBAL memory.
TAIL_HEADER_FIELDS: List[Tuple[str, str]] = [
    ('int', 'magic'), ('int', 'total_bytes'), ('int', 'nk'), ('int', 'n_slots'),
    ('int', 'n_levels'), ('int', 'n_fwd_tiles'), ('int', 'n_bwd_tiles'), ('int', 'n_ops'),
    ('int', 'off_i32'), ('int', 'off_f64'), ('int', 'off_u16'), ('int', 'pad0'),
    ('int', 'i_level_ptr'), ('int', 'i_op_ptr'), ('int', 'i_scale_ptr'), ('int', 'i_tiles'),
    ('int', 'f_S0'), ('int', 'h_rho_slot'), ('int', 'h_level_cols'), ('int', 'h_ops'),
    ('int', 'h_scale'), ('int', 'i_gtgt_ptr'), ('int', 'i_gseg'), ('int', 'h_gops'),     # owner-writes form of the update ops
    ('int', 'i_cround_ptr'), ('int', 'h_cops'), ('int', 'i_cgroup_ptr'), ('int', 'i_clevel_group'),     # coloured-rounds form
]


def tail_header_struct_c(name='CpgTailHeader') -> str:
    return 'struct %s {\n%s};\n' % (name, ''.join(f'  {t} {n};\n' for t, n in TAIL_HEADER_FIELDS))


def pack_tail_blob(T) -> bytes:
    """T: offline.refactor.RefactorTables"""
    ar = _Areas()

    def align_u16(mult):
        while len(ar.u16) % mult:
            ar.u16.append(0)
    hv = dict(magic=MAGIC + 1, nk=T.nk, n_slots=T.n_slots, n_levels=len(T.level_ptr) - 1,
              n_fwd_tiles=len(T.fwd_tiles), n_bwd_tiles=len(T.bwd_tiles), n_ops=len(T.ops))
    hv['i_level_ptr'] = ar.add_i32(T.level_ptr)
    hv['i_op_ptr'] = ar.add_i32(T.op_ptr)
    hv['i_scale_ptr'] = ar.add_i32(T.scale_ptr)
    hv['f_S0'] = ar.add_f64(T.S0)
    hv['h_rho_slot'] = ar.add_u16(T.rho_slot)
    hv['h_level_cols'] = ar.add_u16(T.level_cols)
    align_u16(4)
    hv['h_ops'] = ar.add_u16(T.ops)
    align_u16(2)
    hv['h_scale'] = ar.add_u16(T.scale)
    hv['i_gtgt_ptr'] = ar.add_i32(T.g_tgt_ptr)
    hv['i_gseg'] = ar.add_i32(T.g_seg)
    align_u16(4)
    hv['h_gops'] = ar.add_u16(T.g_ops)
    hv['i_cround_ptr'] = ar.add_i32(T.c_round_ptr)
    hv['i_cgroup_ptr'] = ar.add_i32(T.c_group_ptr)
    hv['i_clevel_group'] = ar.add_i32(T.c_level_group)
    align_u16(4)
    # bit 15 of an op's 4th field (the pivot position j < 32768) = "a warp barrier follows this round" (last round of a sync group)
    cops = np.array(T.c_ops, dtype=np.int64, copy=True).reshape(-1, 4)
    assert T.nk < 32768 and T.n_slots + LANES < 65536
    for g in range(len(T.c_group_ptr) - 1):
        last = int(T.c_group_ptr[g + 1]) - 1
        if last >= int(T.c_group_ptr[g]):
            cops[last * LANES:(last + 1) * LANES, 3] |= 0x8000
    hv['h_cops'] = ar.add_u16(cops)
    # tile header (8 ints): [i32 offset of the packed entries (slot | position << 16, lane-interleaved), 0, K, r_pad, rows,
    #                        u16 offset of the row list, inside: u16 offset of the slot table | first slot of the packed triangle,
    #                        0 no couplings inside / 1 slot table / 2 packed triangle]
    # pad0 = 3: both fields of an entry word are BYTE offsets (index * 8; needs n_slots, nk < 8192), the kernel then spends no
    # shift on them; 0: plain indices (larger families)
    shift = 3 if (T.n_slots < 8192 and T.nk < 8192) else 0
    hv['pad0'] = shift
    tab = []
    for t in list(T.fwd_tiles) + list(T.bwd_tiles):
        words = ((t.slots.astype(np.uint32) << shift) | ((t.cols.astype(np.uint32) << shift) << 16)).astype(np.uint32).view(np.int32)
        hw = ar.add_i32(words)
        hr = ar.add_u16(np.concatenate([t.rows, np.zeros(LANES - len(t.rows), dtype=np.uint16)]))
        if t.inside is None:
            hin, kind = 0, 0
        elif t.dense_base >= 0:
            hin, kind = int(t.dense_base), 2
        else:
            hin, kind = ar.add_u16(t.inside), 1
        tab += [hw, 0, t.slots.shape[0], t.r_pad, len(t.rows), hr, hin, kind]
    while len(ar.i32) % 4:            # tile headers are fetched as two int4
        ar.i32.append(0)
    hv['i_tiles'] = ar.add_i32(tab if tab else [0] * 8)
    fmt = '<' + 'i' * len(TAIL_HEADER_FIELDS)

    def hdr():
        return struct.pack(fmt, *[int(hv.get(n, 0)) for _, n in TAIL_HEADER_FIELDS])
    assert len(hdr()) % 16 == 0

    def pad16(b: bytes) -> bytes:
        return b + b'\0' * ((-len(b)) % 16)
    i32b = pad16(np.asarray(ar.i32, dtype='<i4').tobytes())
    f64b = pad16(np.asarray(ar.f64, dtype='<f8').tobytes())
    u16b = pad16(np.asarray(ar.u16, dtype='<u2').tobytes())
    hv['off_i32'] = len(hdr())
    hv['off_f64'] = hv['off_i32'] + len(i32b)
    hv['off_u16'] = hv['off_f64'] + len(f64b)
    hv['total_bytes'] = hv['off_u16'] + len(u16b)
    return hdr() + i32b + f64b + u16b


# ---------------------------------------------------------------------------------------------------------
# Gradient blob: constants of the batched QP backward pass (csrc/grad_kernel.cuh).  Staged in shared memory,
# except S0 (the regularised KKT values in slot order), which each warp reads once per instance from global.
GRAD_HEADER_FIELDS: List[Tuple[str, str]] = [
    ('int', 'magic'), ('int', 'total_bytes'), ('int', 'n'), ('int', 'm'),
    ('int', 'nk'), ('int', 'npb'), ('int', 'n_prim'), ('int', 'n_slots'),
    ('int', 'off_i32'), ('int', 'off_f64'), ('int', 'off_u16'), ('int', 'n_tent'),
    ('int', 'i_ellA'), ('int', 'i_ellAt'), ('int', 'i_ellP'), ('int', 'i_arow_ptr'),
    ('int', 'i_tptr'), ('int', 'h_pinvx'), ('int', 'h_pinvz'), ('int', 'h_prim'),
    ('int', 'h_arow_slot'), ('int', 'h_tidx'), ('int', 'h_tkind'), ('int', 'f_tval'),
]


def grad_header_struct_c(name='CpgGradHeader') -> str:
    return 'struct %s {\n%s};\n' % (name, ''.join(f'  {t} {n};\n' for t, n in GRAD_HEADER_FIELDS))


def pack_grad_blob(*, n, m, perm, P_upper, A, slot_of, n_slots, Mq_b, Ml_b, Mu_b, npb, prim_idx, reg=1e-6,
                   MP_b=None, MA_b=None):
    """P_upper, A: UNSCALED canonical matrices; slot_of(i, j) -> slot of the lower-triangle entry (pivot positions).
    Returns (blob for shared memory, S0 array for global memory)."""
    nk = n + m
    ar = _Areas()
    pinv = np.empty(nk, dtype=np.int64); pinv[np.asarray(perm)] = np.arange(nk)
    A = sp.csr_matrix(A); A.sort_indices()
    Pfull = sp.csr_matrix(P_upper + sp.triu(P_upper, 1).T)
    # K_reg = [[P + reg I, A'], [A, -reg I]] in slot order  (cvxpygen/writer.py:361-364)
    S0 = np.zeros(n_slots)
    for i in range(n):
        S0[pinv[i]] = Pfull[i, i] + reg
    Pc = sp.coo_matrix(sp.tril(Pfull, -1))
    for i, j, v in zip(Pc.row, Pc.col, Pc.data):
        a, b = pinv[i], pinv[j]
        S0[slot_of(max(a, b), min(a, b))] = v
    arow_ptr, arow_slot = [0], []
    for j in range(m):
        S0[pinv[n + j]] = -reg
        for c, v in zip(A.indices[A.indptr[j]:A.indptr[j + 1]], A.data[A.indptr[j]:A.indptr[j + 1]]):
            a, b = pinv[n + j], pinv[c]
            s_ = slot_of(max(a, b), min(a, b))
            S0[s_] = v
            arow_slot.append(s_)
        arow_ptr.append(len(arow_slot))
    hv = dict(magic=MAGIC + 2, n=n, m=m, nk=nk, npb=npb, n_prim=len(prim_idx), n_slots=n_slots)
    if MP_b is None and MA_b is None:
        hv['i_ellA'] = _add_ell(ar, ell_row_blocks(A, m))
        At_blocks = [(K, v, cidx + n) for K, v, cidx in ell_row_blocks(sp.csr_matrix(A.T), n)]
        hv['i_ellAt'] = _add_ell(ar, At_blocks)
        hv['i_ellP'] = _add_ell(ar, ell_row_blocks(Pfull, n))
    else:       # matrix-parameter family: the products run over the index tables of the matrix blob with per-instance values
        hv['i_ellA'] = hv['i_ellAt'] = hv['i_ellP'] = ar.add_i32([0, 0, 0])
    hv['i_arow_ptr'] = ar.add_i32(arow_ptr)
    hv['h_arow_slot'] = ar.add_u16(arow_slot)
    hv['h_pinvx'] = ar.add_u16(pinv[:n]); hv['h_pinvz'] = ar.add_u16(pinv[n:])
    hv['h_prim'] = ar.add_u16(prim_idx)
    # transposed maps: for every batched parameter entry c the list of (kind, row, coefficient)
    tptr, tidx, tkind, tval = [0], [], [], []
    Ms = [sp.csc_matrix(M) for M in (Mq_b, Ml_b, Mu_b)]
    if MP_b is not None or MA_b is not None:      # kinds 3 / 4: entries of P / A (idx = entry number in CSC order)
        Ms += [sp.csc_matrix(MP_b if MP_b is not None else sp.csr_matrix((0, max(npb, 1)))),
               sp.csc_matrix(MA_b if MA_b is not None else sp.csr_matrix((0, max(npb, 1))))]
    for cidx in range(npb):
        for kind, M in enumerate(Ms):
            s_, e_ = M.indptr[cidx], M.indptr[cidx + 1]
            tidx += M.indices[s_:e_].tolist(); tval += M.data[s_:e_].tolist(); tkind += [kind] * (e_ - s_)
        tptr.append(len(tidx))
    hv['n_tent'] = len(tidx)
    hv['i_tptr'] = ar.add_i32(tptr)
    hv['h_tidx'] = ar.add_u16(tidx); hv['h_tkind'] = ar.add_u16(tkind); hv['f_tval'] = ar.add_f64(tval)
    fmt = '<' + 'i' * len(GRAD_HEADER_FIELDS)

    def hdr():
        return struct.pack(fmt, *[int(hv.get(nm, 0)) for _, nm in GRAD_HEADER_FIELDS])
    assert len(hdr()) % 16 == 0

    def pad16(b: bytes) -> bytes:
        return b + b'\0' * ((-len(b)) % 16)
    i32b = pad16(np.asarray(ar.i32, dtype='<i4').tobytes())
    f64b = pad16(np.asarray(ar.f64, dtype='<f8').tobytes())
    u16b = pad16(np.asarray(ar.u16, dtype='<u2').tobytes())
    hv['off_i32'] = len(hdr()); hv['off_f64'] = hv['off_i32'] + len(i32b); hv['off_u16'] = hv['off_f64'] + len(f64b)
    hv['total_bytes'] = hv['off_u16'] + len(u16b)
    return hdr() + i32b + f64b + u16b, S0


# ---------------------------------------------------------------------------------------------------------
# Matrix-parameter blob: tables of the per-instance osqp_update_data_mat path -- canonicalisation maps
# of the P / A entries, entry coordinates for the in-warp Ruiz equilibration (scale_data, scaling.c:44-156), entry ->
# KKT slot maps for the per-instance assembly (form_KKT / update_KKT_P / update_KKT_A, kkt.c:6-212), and ELL *index*
# tables (entry number + operand position) for the residual products with per-instance values.  Lives in GLOBAL memory.
MAT_HEADER_FIELDS: List[Tuple[str, str]] = [
    ('int', 'magic'), ('int', 'total_bytes'), ('int', 'n'), ('int', 'm'),
    ('int', 'nk'), ('int', 'npb'), ('int', 'nnzP'), ('int', 'nnzA'),
    ('int', 'off_i32'), ('int', 'off_f64'), ('int', 'off_u16'), ('int', 'n_slots'),
    ('int', 'i_ellMP'), ('int', 'i_ellMA'), ('int', 'i_ixA'), ('int', 'i_ixAt'),
    ('int', 'i_ixP'), ('int', 'f_Pbase'), ('int', 'f_Abase'), ('int', 'f_q_un'),
    ('int', 'f_S0'), ('int', 'h_Prow'), ('int', 'h_Pcol'), ('int', 'h_Arow'),
    ('int', 'h_Acol'), ('int', 'h_Pslot'), ('int', 'h_Aslot'), ('int', 'scaling_iters'),
]


def mat_header_struct_c(name='CpgMatHeader') -> str:
    return 'struct %s {\n%s};\n' % (name, ''.join(f'  {t} {n};\n' for t, n in MAT_HEADER_FIELDS))


def _index_ell(M_idx: sp.csr_matrix, n_rows: int, pad_idx: int, col_shift: int = 0):
    """M_idx: CSR whose data = entry number + 1.  Returns ELL blocks (K, idx(K,32), cols(K,32)); padding -> pad_idx."""
    out = []
    for K, vals, cols in ell_row_blocks(M_idx, n_rows):
        idx = np.where(vals > 0, vals - 1, pad_idx).astype(np.int64)
        out.append((K, idx, cols + col_shift))
    return out


def pack_matpar_blob(*, n, m, perm, P_pattern, A_pattern, slot_of, n_slots, sigma, MP_b, MA_b, P_base, A_base,
                     q_un, npb, scaling_iters) -> bytes:
    """P_pattern / A_pattern: (indices, indptr, shape) CSC (P upper triangle).  MP_b / MA_b: CSR maps of the entries
    restricted to the batched-parameter columns; P_base / A_base: entry values with the batched parameters at zero (or, for
    a matrix no batched parameter enters, the unscale_data round trip of the pristine scaled values).  q_un: the linear
    cost that scale_data sees inside osqp_update_P_A (pristine q after unscale_data)."""
    nk = n + m
    ar = _Areas()
    pinv = np.empty(nk, dtype=np.int64); pinv[np.asarray(perm)] = np.arange(nk)
    Pi, Pp, _ = P_pattern
    Ai, Ap, _ = A_pattern
    nnzP, nnzA = len(Pi), len(Ai)
    assert max(nnzP, nnzA) + 1 < 65536
    Prow = np.asarray(Pi, dtype=np.int64); Pcol = np.repeat(np.arange(n), np.diff(Pp))
    Arow = np.asarray(Ai, dtype=np.int64); Acol = np.repeat(np.arange(n), np.diff(Ap))
    assert (Prow <= Pcol).all(), 'P must be stored as its upper triangle'
    Pslot = [int(pinv[r]) if r == c else slot_of(max(pinv[r], pinv[c]), min(pinv[r], pinv[c])) for r, c in zip(Prow, Pcol)]
    Aslot = [slot_of(max(pinv[n + r], pinv[c]), min(pinv[n + r], pinv[c])) for r, c in zip(Arow, Acol)]
    S0 = np.zeros(n_slots)
    S0[pinv[:n]] = sigma
    hv = dict(magic=MAGIC + 3, n=n, m=m, nk=nk, npb=npb, nnzP=nnzP, nnzA=nnzA, n_slots=n_slots, scaling_iters=int(scaling_iters))
    hv['i_ellMP'] = _add_ell(ar, ell_row_blocks(MP_b, nnzP))
    hv['i_ellMA'] = _add_ell(ar, ell_row_blocks(MA_b, nnzA))
    Aidx = sp.csr_matrix(sp.csc_matrix((np.arange(1, nnzA + 1, dtype=float), Ai, Ap), shape=(m, n)))
    hv['i_ixA'] = _add_ixell(ar, _index_ell(Aidx, m, nnzA))
    hv['i_ixAt'] = _add_ixell(ar, _index_ell(sp.csr_matrix(Aidx.T), n, nnzA, col_shift=n))
    Pidx_u = sp.csc_matrix((np.arange(1, nnzP + 1, dtype=float), Pi, Pp), shape=(n, n))
    Pidx = sp.csr_matrix(Pidx_u + sp.triu(Pidx_u, 1).T)
    hv['i_ixP'] = _add_ixell(ar, _index_ell(Pidx, n, nnzP))
    hv['f_Pbase'] = ar.add_f64(P_base); hv['f_Abase'] = ar.add_f64(A_base)
    hv['f_q_un'] = ar.add_f64(q_un); hv['f_S0'] = ar.add_f64(S0)
    hv['h_Prow'] = ar.add_u16(Prow); hv['h_Pcol'] = ar.add_u16(Pcol)
    hv['h_Arow'] = ar.add_u16(Arow); hv['h_Acol'] = ar.add_u16(Acol)
    hv['h_Pslot'] = ar.add_u16(Pslot); hv['h_Aslot'] = ar.add_u16(Aslot)
    fmt = '<' + 'i' * len(MAT_HEADER_FIELDS)

    def hdr():
        return struct.pack(fmt, *[int(hv.get(nm, 0)) for _, nm in MAT_HEADER_FIELDS])
    assert len(hdr()) % 16 == 0

    def pad16(b: bytes) -> bytes:
        return b + b'\0' * ((-len(b)) % 16)
    i32b = pad16(np.asarray(ar.i32, dtype='<i4').tobytes())
    f64b = pad16(np.asarray(ar.f64, dtype='<f8').tobytes())
    u16b = pad16(np.asarray(ar.u16, dtype='<u2').tobytes())
    hv['off_i32'] = len(hdr()); hv['off_f64'] = hv['off_i32'] + len(i32b); hv['off_u16'] = hv['off_f64'] + len(f64b)
    hv['total_bytes'] = hv['off_u16'] + len(u16b)
    return hdr() + i32b + f64b + u16b


def _add_ixell(ar: _Areas, blocks) -> int:
    """int32 table: per 32-row block [K, u16 offset of the entry numbers, u16 offset of the operand positions]."""
    table = []
    for K, idx, cols in blocks:
        table += [K, ar.add_u16(idx), ar.add_u16(cols)]
    return ar.add_i32(table) if table else ar.add_i32([0, 0, 0])
