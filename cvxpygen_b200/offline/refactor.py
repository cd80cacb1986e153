"""Tables for the in-kernel numeric LDL' re-factorisation (the "tail" path).

When OSQP adapts rho (`adapt_rho`, osqp_sources/src/auxil.c:54-74 ->
`osqp_update_rho`, src/osqp.c:1268-1325 -> `update_KKT_param2`, src/kkt.c:214-222 ->
`QDLDL_factor`, qdldl.c:72-233) or when an instance's bounds change a constraint's type
(`update_rho_vec`, auxil.c:100-142), that instance needs its OWN factor of
K(rho_vec) = [[P+sigma I, A'],[A, -diag(1/rho_vec)]].  The symbolic structure (ordering,
elimination tree, pattern of L) is the family's; only values change.  QDLDL does an
up-looking, row-by-row numeric factorisation -- sequential.  For a warp we pre-compute:

  * S layout: value slot of every entry of the lower triangle of P K P' inside the pattern of L
    (slots [0,nk) = diagonal; then, for every solve group whose diagonal block is at least half full, a packed
    strict-lower triangle -- padding included -- that the kernel addresses arithmetically; then the remaining
    strictly-lower entries column by column);
  * S0: slot values of K without the -1/rho diagonal;  rho_slot[j] = diagonal slot of constraint j;
  * per elimination-tree level (columns of a level are independent):
      - the columns of the level                          -> D_j = S[j], S[j] <- 1/D_j
      - update ops (t, a, b, j):  S[t] -= S[a] * S[b] * Dinv_j   (right-looking; atomics resolve
        different columns of the level hitting the same target)
      - scale list (slot):         S[slot] *= Dinv_{col(slot)}   (Y -> L)
  * triangular-solve tiles over per-instance values: like offline/schedule.py tiles but each entry
    carries a SLOT into S instead of an inlined coefficient; one set of tiles per level, rows of thin
    levels are spread over several lanes (p-split) so that a chain level costs ~1 FMA + shuffles.

`emulate()` runs the exact op sequence in numpy; tests pin it against a dense solve.
"""
from dataclasses import dataclass
from typing import List

import numpy as np
import scipy.sparse as sp

from .kkt import LDLFactor
from .schedule import LANES, _pad_pow2


@dataclass
class SlotTile:
    rows: np.ndarray     # (nrows,) pivot positions written
    r_pad: int
    slots: np.ndarray    # (K, 32) uint16 slot into S   (padding: slot of a zero entry, see zero_slot)
    cols: np.ndarray     # (K, 32) uint16 position in w
    dense_base: int = -1        # >= 0: the couplings INSIDE this group tile sit in a packed strict-lower triangle starting at
    #                             this slot (see tri_offset): the kernel addresses them arithmetically, no index table
    inside: np.ndarray = None   # group tiles only: (nrows, 32) uint16, inside[j, t] = slot of the coefficient that
    #                             couples row t to row j of the SAME tile (zero_slot if none).  The kernel resolves these
    #                             dependencies with an in-register sweep (one shuffle + FMA per row) instead of one
    #                             shared-memory round trip per elimination-tree level.


@dataclass
class RefactorTables:
    nk: int
    n_slots: int                  # nk + nnz(L) + 1 (last slot is a constant zero for padding)
    S0: np.ndarray                # (n_slots,) base values
    rho_slot: np.ndarray          # (m,) diagonal slot of constraint row j
    level_ptr: np.ndarray         # (n_levels+1,) into level_cols
    level_cols: np.ndarray        # columns (pivot positions) level by level
    op_ptr: np.ndarray            # (n_levels+1,) into ops
    ops: np.ndarray               # (n_ops, 4) uint16: t, a, b, j
    scale_ptr: np.ndarray         # (n_levels+1,) into scale
    scale: np.ndarray             # (n_scale, 2) uint16: slot, j
    fwd_tiles: List[SlotTile]
    bwd_tiles: List[SlotTile]
    slot: np.ndarray = None       # (nk, nk) slot of the strictly-lower entry (i, j) in pivot positions, -1 outside the pattern
    # owner-writes ("gather") form of the update ops: per level the distinct targets, most loaded first; target ti owns the
    # ops g_ops[g_seg[ti]:g_seg[ti+1]] (all with the same t) and is the only writer of S[t] in that level -> no atomics, one
    # fixed summation order.  Kernel path: tail_factor under CPG_TAIL_GATHER_FACTOR.
    g_tgt_ptr: np.ndarray = None  # (n_levels+1,) into the target list
    g_seg: np.ndarray = None      # (n_targets+1,) first op of every target
    g_ops: np.ndarray = None      # (n_ops, 4): t, a, b, j -- the ops of `ops`, level by level, grouped by target
    # "coloured rounds" form: per level the ops dealt to rounds of 32 with pairwise DISTINCT targets (padded with a no-op on the zero
    # slot), so a warp applies a round with plain read-modify-writes and one __syncwarp: no atomics, a fixed order per target
    # (bit-reproducible), lanes balanced.  Kernel path: tail_factor, CPG_TAIL_FACTOR_FORM = 2 (default).
    c_round_ptr: np.ndarray = None   # (n_levels+1,) first round of every level
    c_ops: np.ndarray = None         # (n_rounds * 32, 4): t, a, b, j
    c_group_ptr: np.ndarray = None   # sync groups: rounds between two warp barriers (no target twice inside a group)
    c_level_group: np.ndarray = None  # (n_levels+1,) first group of every level

    def slot_of(self, i: int, j: int) -> int:
        if i == j:
            return int(i)
        s = int(self.slot[i, j])
        if s < 0:
            raise KeyError('entry outside the symbolic pattern')
        return s

    @property
    def zero_slot(self):
        return self.n_slots - 1


DENSE_GROUP_MIN_FILL = 0.5     # a group whose diagonal block is at least this full gets the packed-triangle layout


def tri_offset(j: int, g: int) -> int:
    """Packed strict-lower triangle of a g-row group, stored by COLUMN j (the row being swept) with the dependent rows
    t = j+1 .. g-1 contiguous: coupling (t, j) lives at dense_base + tri_offset(j, g) + t."""
    return j * (g - 1) - (j * (j - 1)) // 2 - (j + 1)


def _slot_tiles(rows, entries, zero_slot) -> List[SlotTile]:
    """rows: positions; entries[i] = list of (slot, col) for row rows[i]."""
    tiles = []
    for t0 in range(0, len(rows), LANES):
        sel = list(range(t0, min(t0 + LANES, len(rows))))
        nr = len(sel)
        r_pad = _pad_pow2(nr)
        p = LANES // r_pad
        kmax = max(1, max(len(entries[i]) for i in sel))
        K = -(-kmax // p)
        slots = np.full((K, LANES), zero_slot, dtype=np.uint16)
        cols = np.zeros((K, LANES), dtype=np.uint16)
        for lane in range(LANES):
            r, part = lane % r_pad, lane // r_pad
            if r >= nr:
                cols[:, lane] = rows[sel[0]]
                continue
            e = entries[sel[r]][part::p]
            cols[:, lane] = rows[sel[r]]
            for k, (s, c) in enumerate(e):
                slots[k, lane] = s
                cols[k, lane] = c
        tiles.append(SlotTile(rows=np.asarray(rows)[sel].astype(np.uint16), r_pad=r_pad, slots=slots, cols=cols))
    return tiles


def _level_groups(level: np.ndarray, lo_level: int):
    """Consecutive levels >= lo_level packed into groups of <= 32 rows; a level wider than 32 rows stays alone."""
    n_levels = int(level.max()) + 1
    groups, cur = [], []
    for lv in range(lo_level, n_levels):
        rows = np.nonzero(level == lv)[0].tolist()
        if len(rows) > LANES:
            if cur:
                groups.append(cur); cur = []
            groups.append(rows)
        elif len(cur) + len(rows) > LANES:
            groups.append(cur); cur = rows
        else:
            cur = cur + rows
    if cur:
        groups.append(cur)
    return groups


def _group_tiles(rows, outside, inside_pairs, zero_slot, dense_base=-1):
    """rows sorted ascending; outside[i] = [(slot, col)] entries outside the group; inside_pairs[(t, j)] = slot coupling
    row index t to row index j (both indices into `rows`).  One tile per <= 32 rows; only single-tile groups carry
    inside couplings (wide levels have none)."""
    tiles = _slot_tiles(np.asarray(rows), outside, zero_slot)
    if inside_pairs:
        assert len(tiles) == 1
        g = len(rows)
        ins = np.full((g, LANES), zero_slot, dtype=np.uint16)
        for (t, j), sl in inside_pairs.items():
            ins[j, t] = sl
            if dense_base >= 0:
                assert sl == dense_base + tri_offset(min(t, j), g) + max(t, j)
        tiles[0].inside = ins
        tiles[0].dense_base = dense_base
    return tiles


def build_refactor_tables(F: LDLFactor, K: sp.csc_matrix, n_var: int) -> RefactorTables:
    nk = K.shape[0]
    m = nk - n_var
    perm = F.perm
    pinv = np.empty(nk, dtype=np.int64); pinv[perm] = np.arange(nk)
    patt = F.Lpattern
    level = F.level
    # the triangular solves work on GROUPS of consecutive elimination-tree levels (<= 32 rows); the forward and the backward
    # solve share one partition (level 0 -- the leaves -- only matters to the backward solve and gets its own groups)
    fwd_groups = _level_groups(level, 1)
    lvl0 = np.nonzero(level == 0)[0].tolist()
    bwd_groups = ([lvl0] if lvl0 else []) + fwd_groups
    # slots of strictly-lower entries: first, for every group whose diagonal block is dense enough, a packed triangle
    # (entries outside the symbolic pattern inside it are padding: they stay zero and cost storage only); then the
    # remaining entries column by column
    slot = -np.ones((nk, nk), dtype=np.int64)
    s = nk
    dense_base = {}
    for gi, rows in enumerate(fwd_groups):
        g = len(rows)
        if g < 2 or g > LANES:
            continue
        nin = int(np.tril(patt[np.ix_(rows, rows)], -1).sum())
        if nin < DENSE_GROUP_MIN_FILL * g * (g - 1) / 2:
            continue
        dense_base[gi] = s
        for j in range(g):
            for t in range(j + 1, g):
                if patt[rows[t], rows[j]]:
                    slot[rows[t], rows[j]] = s + tri_offset(j, g) + t
        s += g * (g - 1) // 2
    for j in range(nk):
        rows = np.nonzero(patt[:, j] & (slot[:, j] < 0))[0]
        slot[rows, j] = np.arange(s, s + len(rows))
        s += len(rows)
    n_slots = s + 1
    assert n_slots < 65536
    Kp = K.toarray()[np.ix_(perm, perm)]
    S0 = np.zeros(n_slots)
    S0[:nk] = np.diag(Kp)
    ii, jj = np.nonzero(np.tril(Kp, -1))
    assert (slot[ii, jj] >= 0).all(), 'KKT entry outside the symbolic pattern'
    S0[slot[ii, jj]] = Kp[ii, jj]
    rho_slot = pinv[n_var:]
    S0[rho_slot] = 0.0
    n_levels = int(level.max()) + 1
    level_ptr, level_cols = [0], []
    op_ptr, ops = [0], []
    scale_ptr, scale = [0], []
    for lv in range(n_levels):
        cols = np.nonzero(level == lv)[0]
        level_cols += cols.tolist()
        level_ptr.append(len(level_cols))
        for j in cols:
            rows = np.nonzero(patt[:, j])[0]
            for a_i, i in enumerate(rows):
                for k in rows[:a_i + 1]:
                    t = i if i == k else slot[i, k]          # diagonal slot = position
                    assert t >= 0
                    ops.append((t, slot[i, j], slot[k, j], j))
                scale.append((slot[i, j], j))
        op_ptr.append(len(ops))
        scale_ptr.append(len(scale))
    # triangular solves: groups of consecutive levels (<= 32 rows) resolve their internal dependencies in registers
    fwd, bwd = [], []
    pos_in = {}
    for gi, rows in enumerate(fwd_groups):
        rset = {r: t for t, r in enumerate(rows)}
        outside, inside = [], {}
        for t, i in enumerate(rows):
            ent = []
            for j in np.nonzero(patt[i, :])[0]:
                if j in rset:
                    inside[(t, rset[j])] = slot[i, j]
                else:
                    ent.append((slot[i, j], j))
            outside.append(ent)
        fwd += _group_tiles(rows, outside, inside if len(rows) <= LANES else {}, n_slots - 1, dense_base.get(gi, -1))
    for bi in range(len(bwd_groups) - 1, -1, -1):
        rows = bwd_groups[bi]
        gi = bi - (1 if lvl0 else 0)                 # index of the same group in fwd_groups (-1: the level-0 group)
        rset = {r: t for t, r in enumerate(rows)}
        outside, inside = [], {}
        for t, i in enumerate(rows):
            ent = []
            for k in np.nonzero(patt[:, i])[0]:
                if k in rset:
                    inside[(t, rset[k])] = slot[k, i]
                else:
                    ent.append((slot[k, i], k))
            outside.append(ent)
        if len(rows) > LANES:
            assert not inside
        bwd += _group_tiles(rows, outside, inside, n_slots - 1, dense_base.get(gi, -1) if gi >= 0 else -1)
    ops_arr = np.asarray(ops, dtype=np.int64).reshape(-1, 4)
    g_tgt_ptr, g_seg, g_ops = [0], [0], []
    for lv in range(n_levels):
        o = ops_arr[op_ptr[lv]:op_ptr[lv + 1]]
        if len(o):
            tg, inv, cnt = np.unique(o[:, 0], return_inverse=True, return_counts=True)
            for ti in np.lexsort((tg, -cnt)):                    # most loaded targets first: strided lanes stay balanced
                sel = o[inv == ti]
                g_ops.append(sel)
                g_seg.append(g_seg[-1] + len(sel))
        g_tgt_ptr.append(len(g_seg) - 1)
    g_ops = np.concatenate(g_ops).reshape(-1, 4) if g_ops else np.zeros((0, 4), dtype=np.int64)
    assert len(g_ops) == len(ops_arr)
    c_round_ptr, c_ops = [0], []
    c_group_ptr, c_level_group = [0], [0]       # group g = rounds c_group_ptr[g] .. c_group_ptr[g+1]; level lv = groups c_level_group[lv] .. [lv+1]
    zs = n_slots - 1
    for lv in range(n_levels):
        o = ops_arr[op_ptr[lv]:op_ptr[lv + 1]]
        rounds, first_open = [], 0                      # [set of targets, list of ops]
        for op in o:
            t = int(op[0])
            for r in range(first_open, len(rounds)):
                if len(rounds[r][1]) < LANES and t not in rounds[r][0]:
                    rounds[r][0].add(t); rounds[r][1].append(op); break
            else:
                rounds.append([{t}, [op]])
            while first_open < len(rounds) and len(rounds[first_open][1]) >= LANES:
                first_open += 1
        # rounds whose targets are all new since the last barrier need none between them: SYNC GROUPS (a level that eliminates one
        # column -- the chain of the MPC families -- is a single group: its targets are pairwise distinct)
        seen = set()
        for tg, lst in rounds:
            if seen & tg:
                c_group_ptr.append(len(c_ops) // LANES); seen = set()
            seen |= tg
            # padding: lane l subtracts S[zs]^3 = 0 from ITS OWN dummy slot n_slots + l (the factor storage has 32 slots of slack):
            # no lane ever writes what another lane touches, so a round needs no predicate and racecheck sees no hazard
            c_ops += [tuple(int(v) for v in op) for op in lst] + [(n_slots + l, zs, zs, zs) for l in range(len(lst), LANES)]
        c_round_ptr.append(c_round_ptr[-1] + len(rounds))
        if len(rounds):
            c_group_ptr.append(len(c_ops) // LANES)
        c_level_group.append(len(c_group_ptr) - 1)
    c_ops = np.asarray(c_ops, dtype=np.int64).reshape(-1, 4)
    return RefactorTables(g_tgt_ptr=np.asarray(g_tgt_ptr), g_seg=np.asarray(g_seg), g_ops=g_ops,
                          c_round_ptr=np.asarray(c_round_ptr), c_ops=c_ops,
                          c_group_ptr=np.asarray(c_group_ptr), c_level_group=np.asarray(c_level_group),
                          nk=nk, n_slots=n_slots, S0=S0, rho_slot=rho_slot.astype(np.int64),
                          level_ptr=np.asarray(level_ptr), level_cols=np.asarray(level_cols),
                          op_ptr=np.asarray(op_ptr), ops=np.asarray(ops, dtype=np.int64).reshape(-1, 4),
                          scale_ptr=np.asarray(scale_ptr), scale=np.asarray(scale, dtype=np.int64).reshape(-1, 2),
                          fwd_tiles=fwd, bwd_tiles=bwd, slot=slot)


def emulate_factor(T: RefactorTables, rho_vec: np.ndarray) -> np.ndarray:
    """numpy emulation of the kernel's numeric factorisation; returns S with Dinv on the diagonal slots."""
    S = T.S0.copy()
    S[T.rho_slot] = -1.0 / rho_vec
    for lv in range(len(T.level_ptr) - 1):
        cols = T.level_cols[T.level_ptr[lv]:T.level_ptr[lv + 1]]
        S[cols] = 1.0 / S[cols]
        o = T.ops[T.op_ptr[lv]:T.op_ptr[lv + 1]]
        if len(o):
            np.subtract.at(S, o[:, 0], S[o[:, 1]] * S[o[:, 2]] * S[o[:, 3]])
        sc = T.scale[T.scale_ptr[lv]:T.scale_ptr[lv + 1]]
        if len(sc):
            S[sc[:, 0]] *= S[sc[:, 1]]
    return S


def emulate_factor_gather(T: RefactorTables, rho_vec: np.ndarray) -> np.ndarray:
    """The owner-writes form (tail_factor under CPG_TAIL_GATHER_FACTOR): every target sums its own ops in table order."""
    S = T.S0.copy()
    S[T.rho_slot] = -1.0 / rho_vec
    for lv in range(len(T.level_ptr) - 1):
        cols = T.level_cols[T.level_ptr[lv]:T.level_ptr[lv + 1]]
        S[cols] = 1.0 / S[cols]
        for ti in range(T.g_tgt_ptr[lv], T.g_tgt_ptr[lv + 1]):
            o = T.g_ops[T.g_seg[ti]:T.g_seg[ti + 1]]
            assert (o[:, 0] == o[0, 0]).all()
            acc = 0.0
            for t_, a, b, j in o:
                acc += S[a] * S[b] * S[j]
            S[o[0, 0]] -= acc
        sc = T.scale[T.scale_ptr[lv]:T.scale_ptr[lv + 1]]
        if len(sc):
            S[sc[:, 0]] *= S[sc[:, 1]]
    return S


def emulate_solve(T: RefactorTables, S: np.ndarray, w: np.ndarray) -> np.ndarray:
    """w in pivot positions -> K^{-1} w (pivot positions), using the slot tiles exactly like the kernel."""
    w = np.array(w, dtype=float, copy=True)

    def outside_acc(tile):
        acc = np.zeros(LANES)
        for k in range(tile.slots.shape[0]):
            acc += S[tile.slots[k].astype(int)] * w[tile.cols[k].astype(int)]
        off = LANES // 2
        while off >= tile.r_pad:
            acc = acc + acc[np.arange(LANES) ^ off]
            off //= 2
        return acc

    def run(tile, ascending):
        g = len(tile.rows)
        r = tile.rows.astype(int)
        val = w[r] - outside_acc(tile)[:g]
        if tile.inside is not None:
            order = range(g) if ascending else range(g - 1, -1, -1)
            tt = np.arange(g)
            for j in order:
                vj = val[j]
                mask = (tt > j) if ascending else (tt < j)
                if tile.dense_base >= 0:      # packed triangle, addressed arithmetically like group_sweep_dense in the kernel
                    addr = np.array([tile.dense_base + tri_offset(min(t, j), g) + max(t, j) if t != j else T.zero_slot
                                     for t in range(g)])
                    coef = np.where(mask, S[addr], 0.0)
                else:
                    coef = S[tile.inside[j, :g].astype(int)]
                val = np.where(mask, val - coef * vj, val)
        w[r] = val
    for t in T.fwd_tiles:
        run(t, True)
    w[:T.nk] *= S[:T.nk]                          # D^{-1}
    for t in T.bwd_tiles:
        run(t, False)
    return w


def emulate_factor_coloured(T: RefactorTables, rho_vec: np.ndarray) -> np.ndarray:
    """The coloured-rounds form (tail_factor, CPG_TAIL_FACTOR_FORM = 2): rounds of 32 ops with pairwise distinct targets."""
    S = T.S0.copy()
    S[T.rho_slot] = -1.0 / rho_vec
    for lv in range(len(T.level_ptr) - 1):
        cols = T.level_cols[T.level_ptr[lv]:T.level_ptr[lv + 1]]
        S[cols] = 1.0 / S[cols]
        for g in range(T.c_level_group[lv], T.c_level_group[lv + 1]):
            o = T.c_ops[T.c_group_ptr[g] * LANES:T.c_group_ptr[g + 1] * LANES]
            real = o[:, 0] < T.n_slots                                          # (padding aims at the lanes' dummy slots beyond S)
            assert len(np.unique(o[real, 0])) == real.sum()                 # distinct targets inside a sync group
            o = o[real]
            S[o[:, 0]] = S[o[:, 0]] - S[o[:, 1]] * S[o[:, 2]] * S[o[:, 3]]      # plain read-modify-writes, the whole group at once
        sc = T.scale[T.scale_ptr[lv]:T.scale_ptr[lv + 1]]
        if len(sc):
            S[sc[:, 0]] *= S[sc[:, 1]]
    return S
