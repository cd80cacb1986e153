"""Warp-level schedule for the KKT solve  K x = b  with  P K P' = L D L'.

The reference runs `QDLDL_Lsolve`, a diagonal scale and `QDLDL_Ltsolve`
(qdldl_sources/src/qdldl.c:236-281) column by column on one core: 2*(n+m)
dependent steps per ADMM iteration (a6 -- the dominant cost of the hot path).
One warp per instance cannot afford a dependent step per column, so the factor is
re-expressed offline as a short sequence of *tiles*, each an independent set of
sparse dot products executed by the 32 lanes of the warp:

    forward  groups g=[a,b):  w_g <- T^{-1} w_g - (T^{-1} L[g,:a]) w_{:a}        T = I + L[g,g]
    trailing block  t=[s,n):  w_t <- S^{-1} (w_t - L[t,:s] w_{:s}),  S = L_tt D_tt L_tt'   (optional, dense)
    backward groups g=[a,b):  w_g <- T^{-T} D_g^{-1} w_g - (T^{-T} L[b:,g]') w_{b:}

A group is a union of consecutive elimination-tree levels.  Where a level is wide
(leaves of the tree) T = I and this is classic level scheduling; where the tree
degenerates into a chain (the banded Schur complement of an MPC problem) merging k
levels -- with the small triangular block inverted explicitly offline -- trades a
little fill for k-fold fewer dependent steps.  Rows of a group are cut into tiles of
<= 32 rows; a tile with r <= 16 rows spreads every row over p = 32/r_pad lanes and
finishes with log2(p) shuffle-adds, so short-and-fat blocks still use the whole warp.
Group boundaries, and the start s of the trailing dense block, are chosen by dynamic
programming on a cost model that counts warp-level loop iterations.

Hazard rule that lets the kernel write a tile's rows immediately (no double
buffering): inside a forward group row i only reads rows j <= i of the group, so
tiles are emitted in DEcreasing row order; backward groups read j >= i and are
emitted in INcreasing order.  Within a tile all lanes read before any lane writes.
Rows whose update is the identity (forward leaves) are dropped.
"""
from dataclasses import dataclass, field
from typing import List, Tuple

import numpy as np

from .kkt import LDLFactor

LANES = 32
TILE_OVERHEAD = 3.0       # header decode + sync + store, in units of one inner-loop iteration
REDUCE_COST = 0.75        # one shuffle-add stage


def _pad_pow2(r: int) -> int:
    p = 1
    while p < r:
        p *= 2
    return p


@dataclass
class Tile:
    rows: np.ndarray      # (nrows,) pivot positions written by lanes 0..nrows-1
    r_pad: int            # power of two >= nrows; lane -> (row = lane % r_pad, part = lane // r_pad)
    cols: np.ndarray      # (K, 32) uint16 column positions (padding points at a harmless address, coeff 0)
    vals: np.ndarray      # (K, 32) float64
    row_cols: list = None # per row: sorted column positions of the non-zero coefficients
    row_vals: list = None # per row: the coefficients

    @property
    def K(self):
        return self.cols.shape[0]

    @property
    def parts(self):
        return LANES // self.r_pad


def _tile_cost(nrows: int, kmax: int) -> float:
    r_pad = _pad_pow2(nrows)
    p = LANES // r_pad
    return -(-kmax // p) + REDUCE_COST * np.log2(p) + TILE_OVERHEAD


def _group_cost(lens_in_emit_order: np.ndarray) -> float:
    c = 0.0
    for t0 in range(0, len(lens_in_emit_order), LANES):
        sel = lens_in_emit_order[t0:t0 + LANES]
        c += _tile_cost(len(sel), int(max(1, sel.max())))
    return c


def _make_tiles(rows: np.ndarray, M: np.ndarray, col_pos: np.ndarray, decreasing: bool) -> List[Tile]:
    """rows: pivot positions (ascending); M: dense operator (len(rows) x len(col_pos))."""
    idx = np.arange(len(rows))
    if decreasing:
        idx = idx[::-1]
    tiles = []
    for t0 in range(0, len(idx), LANES):
        sel = idx[t0:t0 + LANES]
        nr = len(sel)
        r_pad = _pad_pow2(nr)
        p = LANES // r_pad
        nz = [np.nonzero(M[i])[0] for i in sel]
        kmax = max(1, max(len(c) for c in nz))
        K = -(-kmax // p)
        cols = np.zeros((K, LANES), dtype=np.uint16)
        vals = np.zeros((K, LANES))
        for lane in range(LANES):
            r, part = lane % r_pad, lane // r_pad
            if r >= nr:
                cols[:, lane] = rows[sel[0]]
                continue
            c = nz[r][part::p]                      # interleave the row's entries over its p lanes
            cols[:len(c), lane] = col_pos[c]
            vals[:len(c), lane] = M[sel[r], c]
            cols[len(c):, lane] = rows[sel[r]]
        tiles.append(Tile(rows=np.asarray(rows)[sel].astype(np.uint16), r_pad=r_pad, cols=cols, vals=vals,
                          row_cols=[np.asarray(col_pos)[c].astype(np.int64) for c in nz],
                          row_vals=[M[i, c].copy() for i, c in zip(sel, nz)]))
    return tiles


def _forward_group(F: LDLFactor, a: int, b: int):
    g = b - a
    Tinv = np.linalg.solve(np.eye(g) + F.L[a:b, a:b], np.eye(g)) if g > 1 else np.ones((1, 1))
    Tinv[np.triu_indices(g, 1)] = 0.0
    M = np.zeros((g, b))
    M[:, a:b] = Tinv
    if a:
        M[:, :a] = -Tinv @ F.L[a:b, :a]
    ident = np.array([np.count_nonzero(M[i]) == 1 and M[i, a + i] == 1.0 for i in range(g)], dtype=bool)
    return np.arange(a, b)[~ident], M[~ident], np.arange(0, b)


def _backward_group(F: LDLFactor, a: int, b: int):
    n = F.L.shape[0]
    g = b - a
    TinvT = np.linalg.solve((np.eye(g) + F.L[a:b, a:b]).T, np.eye(g)) if g > 1 else np.ones((1, 1))
    TinvT[np.tril_indices(g, -1)] = 0.0
    M = np.zeros((g, n - a))
    M[:, :g] = TinvT / F.D[a:b][None, :]
    if b < n:
        M[:, g:] = -TinvT @ F.L[b:, a:b].T
    return np.arange(a, b), M, np.arange(a, n)


def _trailing_block(F: LDLFactor, s: int):
    n = F.L.shape[0]
    Ltt = np.eye(n - s) + F.L[s:, s:]
    Sinv = np.linalg.inv(Ltt @ np.diag(F.D[s:]) @ Ltt.T)
    Sinv = 0.5 * (Sinv + Sinv.T)
    M = np.zeros((n - s, n))
    M[:, s:] = Sinv
    if s:
        M[:, :s] = -Sinv @ F.L[s:, :s]      # pull the forward-solved leading part into the block's rhs
    return np.arange(s, n), M, np.arange(0, n)


@dataclass
class SolveSchedule:
    tiles: List[Tile]              # in execution order
    n: int
    n_fwd_tiles: int = 0
    n_trailing_tiles: int = 0      # tiles of the dense trailing block: ALL are computed before any is written
    trailing_start: int = -1       # pivot position where the dense trailing block starts (-1: none)
    model_cost: float = 0.0

    @property
    def n_entries(self):
        return sum(t.cols.size for t in self.tiles)

    def apply(self, w: np.ndarray) -> np.ndarray:
        """Host emulation of the kernel's tile executor (w in pivot positions; batch on leading axes)."""
        w = np.array(w, dtype=float, copy=True)
        deferred = []
        for it, t in enumerate(self.tiles):
            acc = np.zeros(w.shape[:-1] + (LANES,))
            for k in range(t.K):
                acc = acc + t.vals[k] * w[..., t.cols[k].astype(int)]
            off = LANES // 2
            while off >= t.r_pad:                           # shuffle-xor reduction over the row's lanes
                acc = acc + acc[..., np.arange(LANES) ^ off]
                off //= 2
            if self.n_fwd_tiles <= it < self.n_fwd_tiles + self.n_trailing_tiles:
                deferred.append((t, acc))
                if it == self.n_fwd_tiles + self.n_trailing_tiles - 1:
                    for td, ad in deferred:
                        w[..., td.rows.astype(int)] = ad[..., :len(td.rows)]
                continue
            w[..., t.rows.astype(int)] = acc[..., :len(t.rows)]
        return w


import os as _os
ENCODED_TILE_OVERHEAD = float(_os.environ.get('CPG_TILE_OVERHEAD', 12.0))     # per tile, in wavefront units: result stores, the shuffle-add stages, the hazard on the next tile


def _encoded_cost(tiles: List['Tile'], nk: int) -> float:
    """Modelled shared-memory wavefronts of a group's tiles in their cheaper device encoding (the quantity the kernel is bound by)."""
    c = 0.0
    for t in tiles:
        sp_ = encode_sparse(t, conflict_aware=False)       # the conflict-aware deal removes ~3/4 of the counted conflicts
        de = encode_dense(t)
        ok = max(c0 + (LANES // t.r_pad) * Kp for c0, Kp in de['segs']) <= nk
        c += min(sp_['K'] * SPARSE_STEP_WF + 0.25 * sp_['conflicts'], de['cost'] if ok else float('inf')) + ENCODED_TILE_OVERHEAD
    return c


def build_schedule(F: LDLFactor, max_group_rows: int = 64, allow_trailing: bool = True, cost: str = 'encoded') -> SolveSchedule:
    """Group boundaries by dynamic programming.  cost = 'encoded': a group costs the modelled wavefronts of its tiles as the kernel
    will execute them (round 2; MPC-12/4/10: 1 760 -> 1 435 wavefronts per solve against the step-count model of round 1, which
    remains available as cost = 'count')."""
    level = F.level
    n = len(level)
    bounds = [0] + [k for k in range(1, n) if level[k] != level[k - 1]] + [n]
    nb = len(bounds)
    if cost == 'encoded' and nb * min(nb, max_group_rows) > 20000:
        cost = 'count'                                     # very deep elimination trees: keep the set-up time bounded

    def cost_f(i, j):
        rows, M, cp = _forward_group(F, bounds[i], bounds[j])
        if not len(rows):
            return 0.0
        if cost == 'encoded':
            return _encoded_cost(_make_tiles(rows, M, cp, decreasing=True), n)
        return _group_cost(np.count_nonzero(M, axis=1)[::-1])

    def cost_b(i, j):
        rows, M, cp = _backward_group(F, bounds[i], bounds[j])
        if cost == 'encoded':
            return _encoded_cost(_make_tiles(rows, M, cp, decreasing=False), n)
        return _group_cost(np.count_nonzero(M, axis=1))

    INF = float('inf')
    Ff, Pf = [INF] * nb, [-1] * nb
    Bb, Pb = [INF] * nb, [-1] * nb
    Ff[0] = Bb[0] = 0.0
    for j in range(1, nb):
        for i in range(j - 1, -1, -1):
            if bounds[j] - bounds[i] > max_group_rows and i != j - 1:
                break
            cf = Ff[i] + cost_f(i, j)
            if cf < Ff[j]:
                Ff[j], Pf[j] = cf, i
            cb = Bb[i] + cost_b(i, j)
            if cb < Bb[j]:
                Bb[j], Pb[j] = cb, i
    best_j, best = nb - 1, Ff[nb - 1] + Bb[nb - 1]
    if allow_trailing:
        for j in range(1, nb - 1):
            r = n - bounds[j]
            if r > 160:
                continue
            rt, Mt, cpt = _trailing_block(F, bounds[j])
            dense = (_encoded_cost(_make_tiles(rt, Mt, cpt, decreasing=False), n) if cost == 'encoded'
                     else _group_cost(np.count_nonzero(Mt, axis=1)))
            tot = Ff[j] + Bb[j] + dense
            if tot < best:
                best, best_j = tot, j

    def backtrack(prev, j):
        out = []
        while j > 0:
            out.append((bounds[prev[j]], bounds[j]))
            j = prev[j]
        return out[::-1]

    tiles: List[Tile] = []
    for a, b in backtrack(Pf, best_j):
        rows, M, cp = _forward_group(F, a, b)
        if len(rows):
            tiles += _make_tiles(rows, M, cp, decreasing=True)
    n_fwd = len(tiles)
    s = -1
    if best_j != nb - 1:
        s = bounds[best_j]
        rows, M, cp = _trailing_block(F, s)
        # the trailing block reads every row of the block: emit as ONE hazard unit -> handled by the
        # kernel as a multi-tile "gather-then-write" step (flagged through SolveSchedule.trailing_start)
        tiles += _make_tiles(rows, M, cp, decreasing=False)
    n_trail = len(tiles) - n_fwd
    for a, b in reversed(backtrack(Pb, best_j)):
        rows, M, cp = _backward_group(F, a, b)
        tiles += _make_tiles(rows, M, cp, decreasing=False)
    return SolveSchedule(tiles=tiles, n=n, n_fwd_tiles=n_fwd, n_trailing_tiles=n_trail,
                         trailing_start=s, model_cost=best)


# ---------------------------------------------------------------------------------------------------------
# Device encodings of a tile for the two-instances-per-warp kernel (admm_pair_kernel.cuh)
#   sparse: lane-interleaved ELL, column indices packed two per 32-bit word; every lane gathers its own w entry
#   dense : the union of the rows' columns is covered by a few contiguous column segments; all lanes of a
#           row-part read the SAME w address at each step (shared-memory broadcast), no index loads at all.
# Shared-memory wavefronts per inner step, MEASURED per instruction on B200 (ncu source page of the MPC kernel, round 2,
# profiles/r2_multi_v8_ncu_summary.md): a coefficient LDS.64 is 2 wavefronts, a broadcast LDS.128 of the operand pair is 2 (not 1),
# a 16-byte gather is 4 plus its bank conflicts, the packed index word 0.5.  (Round 1 assumed 9 / 3, which encoded four tiles of
# the MPC schedule dense although their gather form is cheaper: 1 760 -> 1 618 modelled wavefronts per solve.)
SPARSE_STEP_WF = float(__import__('os').environ.get('CPG_SPARSE_WF', 6.5))      # (environment overrides: A/B sweeps only)
DENSE_STEP_WF = 4.0
DENSE_SEG_OVERHEAD = 0.5


def _quarter_conflicts(cols_k: np.ndarray) -> int:
    """Extra shared-memory wavefronts of one 16-byte gather step: a quarter-warp (8 lanes x 16 B = all 32 banks once) is served in
    one pass iff its DISTINCT positions fall into distinct bank groups (position mod 8); equal positions are a broadcast."""
    extra = 0
    for q in range(0, LANES, 8):
        pos = np.unique(cols_k[q:q + 8])
        extra += int(np.bincount(pos % 8, minlength=8).max()) - 1
    return extra


def encode_sparse(t: Tile, conflict_aware: bool = True):
    """Lane-interleaved ELL.  The ORDER in which a lane walks its entries is free (a dot product), so the entries are dealt to the
    steps such that the eight lanes of every quarter-warp hit distinct bank groups of the interleaved pair vector wherever the
    tile allows it (ncu, round 1: 9.5 % of the main kernel's shared-memory wavefronts were conflicts of these gathers)."""
    p = LANES // t.r_pad
    nr = len(t.rows)
    kmax = max(1, max(len(c) for c in t.row_cols))
    K = -(-kmax // p)
    K += K % 2                                     # pairs of steps share one packed index word
    cols = np.zeros((K, LANES), dtype=np.int64)
    vals = np.zeros((K, LANES))
    ent = []                                       # per lane: list of (col, val)
    for lane in range(LANES):
        r, part = lane % t.r_pad, lane // t.r_pad
        if r >= nr:
            ent.append([]); continue
        ent.append(list(zip([int(c) for c in t.row_cols[r][part::p]], [float(v) for v in t.row_vals[r][part::p]])))
    if not conflict_aware:
        for lane in range(LANES):
            for k, (c, v) in enumerate(ent[lane]):
                cols[k, lane] = c; vals[k, lane] = v
    else:
        rem = [list(e) for e in ent]
        for k in range(K):
            left = K - k
            for q in range(0, LANES, 8):
                lanes = sorted(range(q, q + 8), key=lambda l: -len(rem[l]))
                taken = {}                          # bank group -> position already read by this quarter in this step
                for l in lanes:
                    if not rem[l]:
                        continue
                    must = len(rem[l]) >= left      # no slack left: this lane has to place an entry now
                    pick = None
                    for i, (c, v) in enumerate(rem[l]):
                        g = c % 8
                        if g not in taken or taken[g] == c:
                            pick = i; break
                    if pick is None:
                        if not must:
                            continue                # wait for a later step (padding now)
                        pick = 0
                    c, v = rem[l].pop(pick)
                    taken.setdefault(c % 8, c)
                    cols[k, l] = c; vals[k, l] = v
        assert not any(rem)
    # padding entries (coefficient 0) copy an address of their own quarter-warp (a broadcast, never a conflict); a quarter without
    # any entry in this step reads the step's first address
    for k in range(K):
        used_all = cols[k][vals[k] != 0]
        for q in range(0, LANES, 8):
            sl = slice(q, q + 8)
            real = vals[k][sl] != 0
            fill = cols[k][sl][real][0] if real.any() else (used_all[0] if len(used_all) else int(t.rows[0]))
            cq = cols[k][sl]; cq[~real] = fill
    packed = (cols[0::2] | (cols[1::2] << 16)).astype(np.uint32)
    conflicts = sum(_quarter_conflicts(cols[k]) for k in range(K))
    return dict(kind=0, K=K, vals=vals, cols32=packed, cost=K * SPARSE_STEP_WF + conflicts, conflicts=conflicts)


def encode_dense(t: Tile, max_gap: int = 3):
    p = LANES // t.r_pad
    nr = len(t.rows)
    used = np.unique(np.concatenate([c for c in t.row_cols if len(c)])) if any(len(c) for c in t.row_cols) else np.array([int(t.rows[0])])
    segs = []                                      # (c0, width)
    c0 = prev = int(used[0])
    for c in used[1:]:
        c = int(c)
        if c - prev > max_gap * p:
            segs.append((c0, prev - c0 + 1)); c0 = c
        prev = c
    segs.append((c0, prev - c0 + 1))
    dense = {}
    for r in range(nr):
        for c, v in zip(t.row_cols[r], t.row_vals[r]):
            dense[(r, int(c))] = v
    seg_tab, blocks, steps = [], [], 0
    for c0, W in segs:
        Kp = -(-W // p)
        vals = np.zeros((Kp, LANES))
        for lane in range(LANES):
            r, part = lane % t.r_pad, lane // t.r_pad
            if r >= nr:
                continue
            for k in range(Kp):
                vals[k, lane] = dense.get((r, c0 + part + p * k), 0.0)
        seg_tab.append((c0, Kp)); blocks.append(vals); steps += Kp
    return dict(kind=1, segs=seg_tab, vals=np.concatenate(blocks, axis=0), K=steps,
                cost=steps * DENSE_STEP_WF + len(segs) * DENSE_SEG_OVERHEAD, n_addr=int(used.max()) + p)


def encode_best(t: Tile, nk: int):
    sp_, de = encode_sparse(t), encode_dense(t)
    # a dense segment may read up to p-1 addresses past its last used column: they must exist
    if de['cost'] < sp_['cost'] and max(c0 + (LANES // t.r_pad) * Kp for c0, Kp in de['segs']) <= nk:
        return de
    return sp_


def apply_encoded(sched: 'SolveSchedule', enc: list, w: np.ndarray) -> np.ndarray:
    """Host emulation of the pair kernel's tile executor on the ENCODED tiles (single right-hand side)."""
    w = np.array(w, dtype=float, copy=True)
    deferred = []
    lanes = np.arange(LANES)
    for it, (t, e) in enumerate(zip(sched.tiles, enc)):
        p = LANES // t.r_pad
        acc = np.zeros(LANES)
        if e['kind'] == 0:
            for k in range(e['K']):
                cc = e['cols32'][k // 2]
                c = (cc & 0xffff) if k % 2 == 0 else (cc >> 16)
                acc += e['vals'][k] * w[c.astype(int)]
        else:
            row0 = 0
            part = lanes // t.r_pad
            for c0, Kp in e['segs']:
                for k in range(Kp):
                    acc += e['vals'][row0 + k] * w[c0 + part + p * k]
                row0 += Kp
        off = LANES // 2
        while off >= t.r_pad:
            acc = acc + acc[lanes ^ off]
            off //= 2
        if sched.n_fwd_tiles <= it < sched.n_fwd_tiles + sched.n_trailing_tiles:
            deferred.append((t, acc))
            if it == sched.n_fwd_tiles + sched.n_trailing_tiles - 1:
                for td, ad in deferred:
                    w[td.rows.astype(int)] = ad[:len(td.rows)]
            continue
        w[t.rows.astype(int)] = acc[:len(t.rows)]
    return w
