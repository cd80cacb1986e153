"""Generation-time setup of one SOCP family for the IPM-CUDA backend (BASELINE config 3).

What the reference does once per problem in ECOS_setup (cvxpygen/solvers/ecos/src/preproc.c:600-1000): equilibrate the
data (src/equil.c:210-340), build the 'stretched' KKT skeleton with its static regularisation (src/preproc.c:77-330),
order it with AMD and run the symbolic LDL'.  Here the same steps produce the tables of a CTA-per-instance kernel:

  * everything lives in ONE index space, the natural stretched order  k = [x (n) | y (p) | z stretched (m + 2 nsoc)],
    so the constraint matrix M = [A ; G stretched] is a single COO table (sorted by row) that serves both the residuals
    of the iterate and the refinement residuals of the KKT solves;
  * the elimination order is a minimum-degree order re-sequenced by elimination-tree level (cvxpygen_b200.offline.kkt);
    the leading WIDE levels are processed by all threads with one barrier per level, the trailing chain of narrow
    levels is closed into a dense block of at most 32 columns that one warp factors and solves with shuffles;
  * the numeric factorisation is table driven: a slot array S = [L entries of wide columns | diagonal | dense tail
    block], a base image of the constant KKT entries, and a list of update operations (target, a, b, column) sorted by
    the level of the target's column;  L is kept column-scaled by D (S_ij = L_ij D_j), Dinv separately;
  * triangular solves are pull-form: forward entries sorted by (level of row, row), backward entries in slot order.
"""
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np
import scipy.sparse as sp

from ..ir import CanonFamily
from . import kkt as _kkt
from . import gather as _gather

DELTASTAT = 7e-8          # ecos/include/ecos.h:54
EQUIL_ITERS = 3           # ecos/include/ecos.h:81
MAX_TAIL = 16             # the dense tail block is factored in the registers of one warp (3 x MAX_TAIL doubles per lane)
DEFAULT_THREADS = 256    # threads of the CTA that solves one instance (the gather plans are dealt for this width)


def ecos_equilibrate(A: sp.csc_matrix, G: sp.csc_matrix, l: int, q: List[int], iters: int = EQUIL_ITERS):
    """Ruiz equilibration as ECOS applies it (use_ruiz_equilibration, ecos/src/equil.c:210-340): per pass the scaling
    is sqrt(max |entry|) per row and per column, with one shared value -- the SUM of the row maxima -- for all rows of a
    second-order cone, values below 1e-6 replaced by 1; rows are divided first, then columns."""
    A = sp.csc_matrix(A, dtype=float, copy=True); G = sp.csc_matrix(G, dtype=float, copy=True)
    n, p, m = G.shape[1], A.shape[0], G.shape[0]
    xe, Ae, Ge = np.ones(n), np.ones(p), np.ones(m)
    cone_of = np.arange(m)
    o = l
    for d in q:
        cone_of[o:o + d] = o; o += d
    for _ in range(iters):
        cmax = np.maximum(_kkt_absmax_cols(A, n), _kkt_absmax_cols(G, n))
        ra = _kkt_absmax_rows(A, p)
        rg = _kkt_absmax_rows(G, m)
        tot = np.zeros(m)
        for i in range(m):                      # sequential accumulation, like the C loop
            tot[cone_of[i]] += rg[i]
        rg = tot[cone_of]
        xt, at, gt = (np.where(np.abs(v) < 1e-6, 1.0, np.sqrt(v)) for v in (cmax, ra, rg))
        A = _div_rows_cols(A, at, xt); G = _div_rows_cols(G, gt, xt)
        xe *= xt; Ae *= at; Ge *= gt
    return A, G, xe, Ae, Ge


def _kkt_absmax_cols(M, n):
    out = np.zeros(n)
    if M.nnz:
        np.maximum.at(out, np.repeat(np.arange(n), np.diff(M.indptr)), np.abs(M.data))
    return out


def _kkt_absmax_rows(M, m):
    out = np.zeros(m)
    if M.nnz:
        np.maximum.at(out, M.indices, np.abs(M.data))
    return out


def _div_rows_cols(M, r, c):
    M = M.copy()
    cols = np.repeat(np.arange(M.shape[1]), np.diff(M.indptr))
    M.data = (M.data / r[M.indices]) / c[cols]
    return M


@dataclass
class SOCPSetup:
    family: CanonFamily
    batch_params: List[str]
    n: int
    p: int
    m: int
    l: int
    q: List[int]
    mt: int
    nk: int
    npb: int
    xe: np.ndarray
    Ae: np.ndarray
    Ge: np.ndarray
    A_eq: sp.csc_matrix
    G_eq: sp.csc_matrix
    perm: np.ndarray                 # position -> k
    pos_level: np.ndarray
    n_wide_levels: int
    t0: int                          # first tail position
    tables: Dict[str, np.ndarray] = field(default_factory=dict)
    defines: Dict[str, int] = field(default_factory=dict)
    level_ranges: Dict[str, List[int]] = field(default_factory=dict)
    smem_blob: bytes = b''
    gmem_blob: bytes = b''
    prim_idx: Optional[np.ndarray] = None
    dual_idx: Optional[np.ndarray] = None
    stats: Dict[str, float] = field(default_factory=dict)
    plans: Dict[str, object] = field(default_factory=dict)
    threads: int = DEFAULT_THREADS

    @property
    def nt(self):
        return self.nk - self.t0


def stretch_layout(l: int, q: List[int]):
    """z index -> stretched index; per cone (z offset, stretched offset, size)."""
    m = l + sum(q)
    zmap = np.zeros(m, dtype=np.int64); zmap[:l] = np.arange(l)
    blocks = []
    o, so = l, l
    for d in q:
        zmap[o:o + d] = so + np.arange(d)
        blocks.append((o, so, d)); o += d; so += d + 2
    return zmap, blocks, m + 2 * len(q)


def setup_socp_family(fam: CanonFamily, batch_params: Optional[List[str]] = None,
                      theta: Optional[np.ndarray] = None, threads: int = DEFAULT_THREADS) -> SOCPSetup:
    if fam.solver_type != 'conic':
        raise ValueError('IPM-CUDA handles the conic canonical form only')
    if batch_params is None:
        batch_params = [p.name for p in fam.params if not (fam.changes('A', [p.name]) or fam.changes('G', [p.name]))]
    mat_params = []
    for name in batch_params:
        fam.param(name)
        if fam.changes('A', [name]) or fam.changes('G', [name]):
            mat_params.append(name)      # per-instance G / A values: the kernel canonicalises and EQUILIBRATES them itself (IPM_MATPAR)
    theta = fam.theta_default() if theta is None else np.asarray(theta, dtype=float)
    n, p, m = fam.n_var, fam.n_eq, fam.n_ineq
    l, q = int(fam.cone_dims.get('l', 0)), [int(d) for d in fam.cone_dims.get('q', [])]
    assert l + sum(q) == m
    nsoc = len(q)
    A = fam.canon_matrix('A', theta) if p else sp.csc_matrix((0, n))
    G = fam.canon_matrix('G', theta)
    A_eq, G_eq, xe, Ae, Ge = ecos_equilibrate(A, G, l, q)
    zmap, blocks, mt = stretch_layout(l, q)
    nk = n + p + mt
    zoff = n + p
    # ---- M = [A ; G stretched] in k-space rows
    Ac, Gc = A_eq.tocoo(), G_eq.tocoo()
    rows = np.r_[n + Ac.row, zoff + zmap[Gc.row]].astype(np.int64)
    cols = np.r_[Ac.col, Gc.col].astype(np.int64)
    vals = np.r_[Ac.data, Gc.data]
    order = np.lexsort((cols, rows))           # (tocoo of a CSC matrix lists the stored entries in CSC order: index = data index)
    mr_t, mr_s, ag_val = rows[order], cols[order], vals[order]
    # ---- KKT pattern (k-space), ordering, levels
    pr = np.r_[np.arange(nk), mr_t, mr_s]; pc = np.r_[np.arange(nk), mr_s, mr_t]
    for o, so, d in blocks:
        iv, iu = zoff + so + d, zoff + so + d + 1
        rr = zoff + so + np.arange(d)
        pr = np.r_[pr, rr[1:], np.full(d - 1, iv), rr, np.full(d, iu)]
        pc = np.r_[pc, np.full(d - 1, iv), rr[1:], np.full(d, iu), rr]
    pat = sp.csr_matrix((np.ones(len(pr)), (pr, pc)), shape=(nk, nk))
    md = _kkt.minimum_degree_order(pat)
    struct, parent = _kkt._symbolic(pat, md)
    reseq, _ = _kkt.level_resequence(struct, parent)
    perm = md[reseq]
    struct, parent = _kkt._symbolic(pat, perm)
    _, level = _kkt.level_resequence(struct, parent)
    assert np.all(np.diff(level) >= 0)
    # ---- tail: the longest suffix of whole levels with at most MAX_TAIL columns
    t0 = nk
    for lev in range(int(level.max()), -1, -1):
        first = int(np.searchsorted(level, lev, side='left'))
        if nk - first > MAX_TAIL:
            break
        t0 = first
    nt = nk - t0
    nlw = int(level[t0 - 1]) + 1 if t0 > 0 else 0
    inv = np.empty(nk, dtype=np.int64); inv[perm] = np.arange(nk)
    # ---- slots
    slot_of = {}
    bw_t, bw_s = [], []
    for j in range(t0):
        for i in struct[j]:
            slot_of[(int(i), j)] = len(bw_t)
            bw_t.append(perm[j]); bw_s.append(perm[int(i)])
    NW = len(bw_t)
    DG0 = NW
    TT0 = NW + nk
    NS = TT0 + nt * nt

    def sidx(pi, pj):                   # positions, pi > pj
        if pj < t0:
            return slot_of[(pi, pj)]
        return TT0 + (pi - t0) * nt + (pj - t0)

    def kslot(kr, kc):
        if kr == kc:
            return DG0 + kr
        pi, pj = inv[kr], inv[kc]
        return sidx(max(pi, pj), min(pi, pj))
    # ---- base image of the constant entries
    Sbase = np.zeros(NS)
    Sbase[DG0:DG0 + n] = DELTASTAT
    Sbase[DG0 + n:DG0 + n + p] = -DELTASTAT
    ag_slot = np.array([kslot(int(r), int(c_)) for r, c_ in zip(mr_t, mr_s)], dtype=np.int64)
    assert len(np.unique(ag_slot)) == len(ag_slot)
    if not mat_params:
        Sbase[ag_slot] += ag_val               # (with per-instance matrices the kernel writes these slots from its own values)
    socv, socu = [], []
    for o, so, d in blocks:
        iv, iu = zoff + so + d, zoff + so + d + 1
        socv += [kslot(zoff + so + r, iv) for r in range(1, d)]
        socu += [kslot(zoff + so + r, iu) for r in range(d)]
    # ---- factor operations, pull form, by level of the target's column
    slot_col_start = np.zeros(t0 + 1, dtype=np.int64)
    for j in range(t0):
        slot_col_start[j + 1] = slot_col_start[j] + len(struct[j])
    ops = []
    for j in range(t0):
        R = [int(i) for i in struct[j]]
        base = int(slot_col_start[j])
        kj = int(perm[j])
        for bi, kpos in enumerate(R):
            lvl = int(level[kpos]) if kpos < t0 else nlw
            for ai in range(bi, len(R)):
                ipos = R[ai]
                tgt = DG0 + int(perm[kpos]) if ai == bi else sidx(ipos, kpos)
                ops.append((lvl, tgt, base + ai, base + bi, kj))
    ops.sort()
    ops = np.array(ops, dtype=np.int64).reshape(-1, 5)
    op_lo = [int(np.searchsorted(ops[:, 0], lv, side='left')) for lv in range(nlw + 2)]
    # ---- forward solve entries by (level of row, row)
    fw = []
    for (i, j), s in slot_of.items():
        lvl = int(level[i]) if i < t0 else nlw
        fw.append((lvl, int(perm[i]), int(perm[j]), s))
    fw.sort()
    fw = np.array(fw, dtype=np.int64).reshape(-1, 4)
    fw_lo = [int(np.searchsorted(fw[:, 0], lv, side='left')) for lv in range(nlw + 2)]
    bw_lo = [int(slot_col_start[np.searchsorted(level[:t0], lv, side='left')]) for lv in range(nlw + 1)]
    lev_lo = [int(np.searchsorted(level[:t0], lv, side='left')) for lv in range(nlw + 1)]
    # ---- affine maps of the per-instance vectors, equilibrated, in k-space: cbh = base + Mb theta_b
    bcols = fam.param_columns(batch_params) if batch_params else np.zeros(0, dtype=int)
    npb = len(bcols)
    theta0 = theta.copy(); theta0[bcols] = 0.0
    scale_k = np.ones(nk); scale_k[:n] = 1.0 / xe
    if p:
        scale_k[n:zoff] = 1.0 / Ae
    scale_k[zoff + zmap] = 1.0 / Ge
    scale_cbh = np.ones(nk) if mat_params else scale_k      # per-instance matrices: c, b, h stay RAW, the kernel divides by its scalings
    krow = {'c': np.arange(n), 'b': n + np.arange(p), 'h': zoff + zmap}
    base = np.zeros(nk)
    mt_, mp_, mv_ = [], [], []
    for pid in ('c', 'b', 'h'):
        Mp = fam.maps.get(pid)
        if Mp is None or Mp.shape[0] == 0:
            continue
        base[krow[pid]] = np.asarray(Mp @ theta0).ravel() * scale_cbh[krow[pid]]
        if npb:
            Mb = sp.coo_matrix(Mp[:, bcols])
            mt_ += list(krow[pid][Mb.row]); mp_ += list(Mb.col); mv_ += list(Mb.data * scale_cbh[krow[pid]][Mb.row])
    # ---- per-instance matrix entries (IPM_MATPAR): RAW value of entry e of `ag` = ent_base[e] + sum emap_v * theta_b[emap_p]
    ent_base = np.zeros(len(ag_val)); et_, ep_, ev_ = [], [], []
    if mat_params:
        nA = int(A.nnz) if p else 0
        MA = sp.csr_matrix(fam.maps['A']) if p else sp.csr_matrix((0, len(theta)))
        MG = sp.csr_matrix(fam.maps['G'])
        assert MA.shape[0] == nA and MG.shape[0] == int(G.nnz), 'one row of the entry map per stored entry (CSC order)'
        Mall = sp.vstack([MA, MG]).tocsr()[order]                   # rows in the order of `ag`
        ent_base = np.asarray(Mall @ theta0).ravel()
        Mb = sp.coo_matrix(Mall[:, bcols])
        et_, ep_, ev_ = list(Mb.row), list(Mb.col), list(Mb.data)
    if 'd' in fam.maps and npb and fam.maps['d'][:, bcols].nnz:
        raise ValueError('objective offset d depending on a batched parameter is not generated yet')
    d_const = float(np.asarray(fam.maps['d'] @ theta).ravel()[0]) if 'd' in fam.maps else 0.0
    # ---- output scaling (backscale, ecos/src/ecos.c:1051-1070) and retrieval lists
    unscale = scale_k.copy()                    # x / xe, y / Ae, z / Ge  (then / tau)
    prim_idx = np.concatenate([v.indices for v in fam.variables]) if fam.variables else np.zeros(0, int)
    dual_k = []
    for dv in fam.duals:
        dual_k.append(n + dv.indices if dv.vec == 'y' else zoff + zmap[dv.indices])
    dual_idx = np.concatenate(dual_k) if dual_k else np.zeros(0, int)
    # ---- gather plans (cvxpygen_b200/offline/gather.py): every sparse phase of the kernel in owner-writes form
    assert nk < 2047 and NS < 65535 and len(ag_val) < 65535
    # (a) products with the constant off-diagonal part of K, rows in k order: entry = (index into ag, source k)
    mv_rows = [[] for _ in range(nk)]
    for e, (r, c_) in enumerate(zip(mr_t, mr_s)):
        mv_rows[int(r)].append((int(c_), e)); mv_rows[int(c_)].append((int(r), e))
    plan_mv = _gather.GatherPlan(T=threads, nfields=2, tbits=11, null_entry=(len(ag_val), 0))
    # the kernel runs the second-order cones of a residual on its last warps while the others gather: leave those out
    nwarp = threads // 32
    cone_warps = min(nsoc, nwarp - 1)
    _gather.add_phase(plan_mv, [(k, 0, [(e, src) for src, e in sorted(mv_rows[k])]) for k in range(nk)],
                      use_warps=nwarp - cone_warps)
    # (b) forward substitution: phase 0 scales the leaves, phase lv pulls row k from the columns below, the last phase
    #     (flag 1) collects the tail rows;  entry = (slot of S, source k).  Leaves (level 0) have nothing to pull.
    fw_rows = {}
    for t_, s_, sl in zip(fw[:, 1], fw[:, 2], fw[:, 3]):
        fw_rows.setdefault(int(t_), []).append((int(sl), int(s_)))
    # a null entry multiplies the always-zero slot S[NS] by SOME element of the vector being solved for: it must be one that no
    # thread writes during the phase (compute-sanitizer racecheck flagged the read of u[0] next to its owner's write -- harmless, the
    # product is 0 either way, but a hazard all the same).  Forward phases never write a leaf, backward phases never write a tail row.
    null_fw = int(perm[lev_lo[0]]) if nlw else 0
    null_bw = int(perm[t0]) if nt else int(perm[nk - 1])
    plan_fw = _gather.GatherPlan(T=threads, nfields=2, tbits=11, null_entry=(NS, null_fw))
    _gather.add_phase(plan_fw, [])      # level 0: the leaves are scaled by whoever writes the right-hand side (l0mask)
    for lv in range(1, nlw):
        ks = [int(perm[pos]) for pos in range(lev_lo[lv], lev_lo[lv + 1])]
        _gather.add_phase(plan_fw, [(k, 0, sorted(fw_rows.get(k, []), key=lambda x: inv[x[1]])) for k in ks])
    _gather.add_phase(plan_fw, [(int(k), 1, sorted(fw_rows.get(int(k), []), key=lambda x: inv[x[1]])) for k in perm[t0:]])
    # (c) backward substitution, wide levels from the top: column k pulls from the rows below it
    bw_cols = {}
    for sl, (t_, s_) in enumerate(zip(bw_t, bw_s)):
        bw_cols.setdefault(int(t_), []).append((sl, int(s_)))
    plan_bw = _gather.GatherPlan(T=threads, nfields=2, tbits=11, null_entry=(NS, null_bw))
    for lv in range(nlw - 1, -1, -1):
        ks = [int(perm[pos]) for pos in range(lev_lo[lv], lev_lo[lv + 1])]
        _gather.add_phase(plan_bw, [(k, 0, sorted(bw_cols[k], key=lambda x: inv[x[1]])) for k in ks if k in bw_cols])
    # (d) numeric factorisation: phase lv-1 completes the columns of level lv (flag 1 = a wide column's diagonal: the
    #     pivot is taken at commit), the last phase the dense tail block;  entry = (slot a, slot b, diagonal slot of the
    #     source column -- it holds 1/d once the column is complete --, 0)
    plan_ops = _gather.GatherPlan(T=threads, nfields=4, tbits=16, null_entry=(NS, NS, NS, 0))
    for lv in range(1, nlw + 1):
        o = ops[op_lo[lv]:op_lo[lv + 1]]
        rows_ = {}
        for _, tgt, a_, b_, kj in o:
            rows_.setdefault(int(tgt), []).append((int(a_), int(b_), DG0 + int(kj), 0))
        rl = []
        for tgt in sorted(rows_):
            is_diag = DG0 <= tgt < DG0 + nk
            wide = is_diag and inv[tgt - DG0] < t0
            rl.append((tgt, 1 if wide else 0, rows_[tgt]))
        if lv < nlw:                    # every column of a level >= 1 has a child, hence an update of its diagonal
            have = {t_ for t_, f_, _ in rl if f_}
            assert have == {DG0 + int(perm[pos]) for pos in range(lev_lo[lv], lev_lo[lv + 1])}
        _gather.add_phase(plan_ops, rl)
    l0mask = np.zeros((nk + 15) // 16, dtype=np.int64)          # bit k: column k is a leaf of the elimination tree
    for pos in range(lev_lo[0], lev_lo[1] if nlw else 0):
        l0mask[int(perm[pos]) >> 4] |= 1 << (int(perm[pos]) & 15)
    T = dict(l0mask=l0mask, mr_t=mr_t, mr_s=mr_s, ag_val=ag_val, fw_t=fw[:, 1], fw_s=fw[:, 2], fw_slot=fw[:, 3],
             bw_t=np.array(bw_t, dtype=np.int64), bw_s=np.array(bw_s, dtype=np.int64), socv=np.array(socv, dtype=np.int64),
             socu=np.array(socu, dtype=np.int64), Sbase=Sbase, ops=ops[:, 1:], tail_k=perm[t0:], cbh_base=base,
             map_t=np.array(mt_, dtype=np.int64), map_p=np.array(mp_, dtype=np.int64), map_v=np.array(mv_, dtype=float),
             unscale=unscale, prim_idx=prim_idx, dual_idx=dual_idx, perm=perm,
             ent_base=ent_base, emap_t=np.array(et_, dtype=np.int64), emap_p=np.array(ep_, dtype=np.int64),
             emap_v=np.array(ev_, dtype=float), ag_slot=ag_slot)
    LR = dict(op_lo=op_lo, fw_lo=fw_lo, bw_lo=bw_lo, lev_lo=lev_lo)
    D = dict(N=n, P=p, M=m, L=l, NSOC=nsoc, MT=mt, NK=nk, ZOFF=zoff, NW=NW, DG0=DG0, TT0=TT0, NS=NS, NT=nt, NLW=nlw,
             NNZM=len(ag_val), NOPS=len(ops), NFW=len(fw), NPB=npb, NMAP=len(mv_), NPRIM=len(prim_idx),
             NDUAL=len(dual_idx), QTOT=sum(d - 1 for d in q), IS_MAX=int(fam.is_maximization), NNZA=int(A_eq.nnz),
             THREADS=threads, MATPAR=int(bool(mat_params)), NEMAP=len(ev_))
    st = SOCPSetup(family=fam, batch_params=list(batch_params), n=n, p=p, m=m, l=l, q=q, mt=mt, nk=nk, npb=npb, xe=xe,
                   Ae=Ae, Ge=Ge, A_eq=A_eq, G_eq=G_eq, perm=perm, pos_level=level, n_wide_levels=nlw, t0=t0, tables=T,
                   defines=D, level_ranges=LR, prim_idx=prim_idx, dual_idx=dual_idx,
                   plans=dict(mv=plan_mv, fw=plan_fw, bw=plan_bw, ops=plan_ops), threads=threads)
    st.stats = dict(nnz_L=NW + nt * (nt - 1) // 2, n_levels=int(level.max()) + 1, n_wide_levels=nlw, tail=nt,
                    factor_ops=len(ops), d_const=d_const)
    _pack(st, d_const)
    return st


# ----------------------------------------------------------------------------------------------------------------------
def _pack(st: SOCPSetup, d_const: float):
    """Two byte strings + the offsets the kernel needs (emitted as #defines into the family header):
       smem blob  = [f64: ag_val, 0] [u32: entries of the mv, fw, bw plans] [u16: their descriptors, socv socu tail_k perm]
       gmem blob  = [f64: Sbase cbh_base unscale map_v] [i32: map_t map_p prim_idx dual_idx]
                    [u64: entries of the factorisation plan] [u32: its descriptors]                (staged per CTA / read in place)"""
    T, D, PL = st.tables, st.defines, st.plans

    def cat(parts, dtype):
        offs, chunks, pos = {}, [], 0
        for name, arr in parts:
            a = np.ascontiguousarray(np.asarray(arr).astype(dtype))
            offs[name] = pos
            chunks.append(a); pos += a.size
        return offs, (np.concatenate(chunks) if chunks else np.zeros(0, dtype))

    def pack_entries(plan):
        e = plan.entry_array().astype(np.uint64) * np.uint64(8)        # the kernel wants byte offsets into f64 arrays
        assert e.max(initial=0) < 65536
        w = np.zeros(e.shape[0], dtype=np.uint64)
        for f in range(plan.nfields):
            w |= e[:, f] << np.uint64(16 * f)
        return w
    # shared-memory blob
    f64 = np.r_[T['ag_val'], 0.0]
    eo, e32 = cat([(nm, pack_entries(PL[nm])) for nm in ('mv', 'fw', 'bw')], np.uint32)
    names16 = ['socv', 'socu', 'tail_k', 'perm', 'l0mask']
    for nm in names16:
        a = T[nm]
        assert a.size == 0 or (a.min() >= 0 and a.max() < 65536), nm
    ho, u16 = cat([('mv_d', PL['mv'].desc_array()), ('fw_d', PL['fw'].desc_array()), ('bw_d', PL['bw'].desc_array())] +
                  [(nm, T[nm]) for nm in names16], np.uint16)
    sm = f64.tobytes()
    u32_off = len(sm)
    sm += e32.tobytes()
    u16_off = len(sm)
    sm += u16.tobytes()
    sm += b'\0' * ((-len(sm)) % 16)
    D['SB_BYTES'] = len(sm); D['SB_U32_OFF'] = u32_off; D['SB_U16_OFF'] = u16_off
    for nm in ('mv', 'fw', 'bw'):
        D['E_' + nm.upper()] = eo[nm]
    for nm, o in ho.items():
        D['H_' + nm.upper()] = o
    # global blob
    go, g64 = cat([('Sbase', T['Sbase']), ('cbh_base', T['cbh_base']), ('unscale', T['unscale']), ('map_v', T['map_v']),
                   ('ent_base', T['ent_base']), ('emap_v', T['emap_v'])], np.float64)
    io, i32 = cat([('map_t', T['map_t']), ('map_p', T['map_p']), ('prim_idx', T['prim_idx']), ('dual_idx', T['dual_idx']),
                   ('emap_t', T['emap_t']), ('emap_p', T['emap_p']), ('mr_t', T['mr_t']), ('mr_s', T['mr_s']),
                   ('ag_slot', T['ag_slot'])], np.int32)
    gm = g64.tobytes()
    i32_off = len(gm)
    gm += i32.tobytes()
    gm += b'\0' * ((-len(gm)) % 16)
    ops_off = len(gm)
    gm += pack_entries(PL['ops']).astype(np.uint64).tobytes()
    opd_off = len(gm)
    gm += PL['ops'].desc_array().astype(np.uint32).tobytes()
    gm += b'\0' * ((-len(gm)) % 16)
    D['GB_BYTES'] = len(gm); D['GB_I32_OFF'] = i32_off; D['GB_OPS_OFF'] = ops_off; D['GB_OPD_OFF'] = opd_off
    for nm, o in go.items():
        D['G_' + nm.upper()] = o
    for nm, o in io.items():
        D['GI_' + nm.upper()] = o
    st.smem_blob, st.gmem_blob = sm, gm
    st.stats['smem_blob_bytes'] = len(sm); st.stats['gmem_blob_bytes'] = len(gm)
    for nm, pl in PL.items():
        st.stats[f'rounds_{nm}'] = pl.n_rounds
        st.stats[f'pad_{nm}'] = len(pl.entries)


# ----------------------------------------------------------------------------------------------------------------------
# numpy emulation of the table-driven factorisation and solve (used by the CPU tests to pin the tables)
def fill_slots(st: SOCPSetup, zdiag: np.ndarray, socv_vals=None, socu_vals=None) -> np.ndarray:
    """S image for a KKT whose (3,3) block has diagonal `zdiag` (length mt, k-space order) and the given v / u columns."""
    T, D = st.tables, st.defines
    S = T['Sbase'].copy()
    S[D['DG0'] + D['ZOFF']:D['DG0'] + D['NK']] += zdiag
    if socv_vals is not None and len(T['socv']):
        S[T['socv']] = socv_vals
    if socu_vals is not None and len(T['socu']):
        S[T['socu']] = socu_vals
    return S


def _inv_pivot(d, sg, eps, delta):
    return 1.0 / (sg * delta if sg * d <= eps else d)


def emulate_factor(st: SOCPSetup, S: np.ndarray, sign: np.ndarray, eps=1e-13, delta=2e-7):
    """The kernel's numeric factorisation run from the gather plans: returns the slot array (one extra zero slot; inverse
    pivots in the diagonal slots, inv(L_tail) in the tail block) and, for convenience, the inverse pivots in k order."""
    D, LR, plan = st.defines, st.level_ranges, st.plans['ops']
    DG0, TT0, nlw, nt = D['DG0'], D['TT0'], D['NLW'], D['NT']
    S = np.r_[S, 0.0]
    for k in st.perm[LR['lev_lo'][0]:LR['lev_lo'][1]]:
        S[DG0 + k] = _inv_pivot(S[DG0 + k], sign[k], eps, delta)

    def commit(t, flag, acc):
        v = S[t] - acc
        S[t] = _inv_pivot(v, sign[t - DG0], eps, delta) if flag else v
    for lv in range(1, nlw + 1):
        _gather.run_phase(plan, lv - 1, lambda e: S[e[0]] * S[e[1]] * S[e[2]], commit)
    tk = st.tables['tail_k']
    B = S[TT0:TT0 + nt * nt].reshape(nt, nt)
    dv = np.zeros(nt)
    for j in range(nt):
        dv[j] = _inv_pivot(S[DG0 + tk[j]], sign[tk[j]], eps, delta)
        for i in range(j + 1, nt):
            sij = B[i, j]
            for k in range(j + 1, i):
                B[i, k] -= sij * B[k, j] * dv[j]
            S[DG0 + tk[i]] -= sij * sij * dv[j]
    Lt = np.tril(B, -1) * dv[None, :] + np.eye(nt)
    X = np.linalg.inv(Lt) if nt else Lt
    S[DG0 + tk] = dv
    B[np.tril_indices(nt, -1)] = X[np.tril_indices(nt, -1)]
    S[TT0:TT0 + nt * nt] = B.ravel()
    return S, S[DG0:DG0 + D['NK']].copy()


def emulate_solve(st: SOCPSetup, S: np.ndarray, Dinv: np.ndarray, rhs: np.ndarray) -> np.ndarray:
    """The kernel's ldl_solve from the gather plans (forward with the D-solve folded in, tail by inv(L_tail), backward)."""
    D = st.defines
    DG0, TT0, nlw, nt = D['DG0'], D['TT0'], D['NLW'], D['NT']
    u = rhs.astype(float).copy()

    def fw_commit(k, tail, acc):
        v = u[k] - acc
        u[k] = v if tail else v * S[DG0 + k]
    leaves = st.perm[st.level_ranges['lev_lo'][0]:st.level_ranges['lev_lo'][1]] if nlw else []
    u[leaves] *= S[DG0 + np.asarray(leaves, dtype=int)]
    for lv in range(1, nlw + 1):
        _gather.run_phase(st.plans['fw'], lv, lambda e: S[e[0]] * u[e[1]], fw_commit)
    tk = st.tables['tail_k']
    X = np.tril(S[TT0:TT0 + nt * nt].reshape(nt, nt), -1) + np.eye(nt)
    w = (X @ u[tk]) * S[DG0 + tk]
    u[tk] = X.T @ w

    def bw_commit(k, _, acc):
        u[k] -= S[DG0 + k] * acc
    for lv in range(nlw):
        _gather.run_phase(st.plans['bw'], lv, lambda e: S[e[0]] * u[e[1]], bw_commit)
    return u
