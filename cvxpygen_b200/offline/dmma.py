"""KKT-solve schedule for the FP64 tensor-core kernel (admm_dmma_kernel.cuh): `mma.sync.m8n8k4.f64` with INSTANCES as N.

Why: the warp-per-two-instances kernel is bound by shared-memory bandwidth (78 % of the LSU data pipe, FP64 pipe at 18 %;
profiles/r1_v7_ncu_summary.md) because every coefficient it loads feeds two FMAs.  The KKT factor is the same for every
instance of the batch, so the solve `w <- M w` of a tile is a (rows x cols) x (cols x instances) product: with eight
instances on the N dimension of DMMA.8x8x4 one 8x4 coefficient fragment feeds 256 FMAs and one operand fragment (4 positions x
8 instances) is shared by the 8 rows.  profiles/r2_dmma_probe.jsonl: DMMA reaches the FP64 pipe's peak (64 FMA/clk/SM, the
same as DFMA) from 4 warps and ~54 FMA/clk/SM with both fragments read from shared memory.

Unit of work: a GROUP of four warps owns eight instances.  The tile sequence is the level-group decomposition of
offline/schedule.py (forward groups `w_g <- T^-1 w_g - (T^-1 L[g,:a]) w_:a`, backward groups with D^-1 folded in), re-cut for
this cost model; every tile is a dense operator M (rows x columns) applied out of place:

  rows      -> ROW BLOCKS of 8 consecutive tile rows (the M dimension of the MMA);
  columns   -> per row block, the sorted union of its non-zero columns cut into K-GROUPS of 4 positions (K dimension; the four
               positions are arbitrary -- every lane gathers its own operand -- so sparse tiles cost ceil(|union| / 4) MMAs);
  ITEM      = one (row block, K-group): an 8x4 coefficient block stored COMPRESSED (32-bit occupancy mask + its non-zeros in
               lane order: the blocks are 37 % full on the MPC family) and the 4 operand positions;
  UNIT      = the items of one row block, or of one of P in {1, 2, 4} interleaved parts of it when a tile has fewer row blocks
               than warps; units are dealt to the four warps longest first, lengths padded per round with null items;
  COMMIT JOB= one row block: sums the <= 4 partial 8x8 results its units left in the staging buffer and writes the rows.

Per tile the kernel runs: every warp its units (MMA accumulators -> staging), group barrier, every warp its commit jobs
(staging -> w), group barrier.  `DmmaSchedule.apply` is the numpy restatement of exactly that, used by the CPU tests.
"""
from dataclasses import dataclass, field
from typing import List

import numpy as np

from .kkt import LDLFactor
from . import schedule as _sched

RB, KB, NWARP = 8, 4, 4
MAX_SLOTS = 16            # staging slots (8x8 partial results) per group
TILE_OVERHEAD = 1.0       # two group barriers + staging round trip, in units of one MMA issue slot per warp
REDUCE_OVERHEAD = 0.5     # extra partials to sum per commit job when a row block is split


@dataclass
class Item:
    pos: np.ndarray       # (4,) operand positions (padding repeats a position, coefficient 0)
    A: np.ndarray         # (8, 4) coefficients


@dataclass
class Unit:
    rb: int               # row block of the tile
    part: int
    items: List[Item]
    warp: int = -1
    rnd: int = -1         # round (u index) in the warp's unit list
    slot: int = -1        # staging slot


@dataclass
class DmmaTile:
    rows: np.ndarray      # positions written (ascending)
    parts: int
    units: List[Unit]
    round_len: List[int] = field(default_factory=list)     # padded items per round
    n_rb: int = 0


def w8_row_offset(p: int) -> int:
    """byte offset of instance column 0 of position p in the group's interleaved work vectors, swizzle bits included: the double
    of instance c sits at `w8_row_offset(p) ^ (c << 3)` (kernel: w8_off)."""
    return (p << 6) | ((((p >> 2) & 3) << 1) << 3)


def _row_blocks(rows, M, col_pos):
    """per 8-row block: (row positions, list of K-groups (pos[4], A[8,4]))"""
    col_pos = np.asarray(col_pos)
    out = []
    for r0 in range(0, len(rows), RB):
        sub = M[r0:r0 + RB]
        used = np.nonzero(np.any(sub != 0, axis=0))[0]
        # a position's 64-byte row of w8 covers banks 0-15 (even position) or 16-31 (odd): two of each per K-group when possible
        ev, od = [c for c in used if col_pos[c] % 2 == 0], [c for c in used if col_pos[c] % 2 == 1]
        mixed = []
        while ev or od:
            for src in (ev, od, ev, od):
                if src:
                    mixed.append(src.pop(0))
                elif ev or od:
                    mixed.append((ev or od).pop(0))
        used = np.asarray(mixed[:len(used)], dtype=np.int64) if len(used) else used
        groups = []
        for k0 in range(0, len(used), KB):
            cols = used[k0:k0 + KB]
            A = np.zeros((RB, KB)); A[:sub.shape[0], :len(cols)] = sub[:, cols]
            pos = np.full(KB, int(col_pos[cols[-1]]), dtype=np.int64); pos[:len(cols)] = col_pos[cols]
            groups.append(Item(pos, A))
        if not groups:       # an all-zero row block still has to write zeros
            groups.append(Item(np.full(KB, int(rows[r0]), dtype=np.int64), np.zeros((RB, KB))))
        out.append((np.asarray(rows[r0:r0 + RB]), groups))
    return out


def _deal(lengths):
    """units longest first, dealt round-robin to the 4 warps; returns (order, round lengths)"""
    order = np.argsort(-np.asarray(lengths), kind='stable')
    rounds = [int(lengths[order[k]]) for k in range(0, len(order), NWARP)]
    rounds = [r + (r % 2) for r in rounds]            # two accumulator chains per unit: even item counts
    return order, rounds


def _tile_cost(n_groups_per_rb):
    best = None
    for P in (1, 2, 4):
        lens = [len(range(p, k, P)) for k in n_groups_per_rb for p in range(P)]
        if len(lens) > MAX_SLOTS:
            continue
        _, rounds = _deal(lens)
        c = sum(rounds) + TILE_OVERHEAD + (REDUCE_OVERHEAD * (P - 1) * -(-len(n_groups_per_rb) // NWARP))
        if best is None or c < best[0]:
            best = (c, P)
    return best


def _make_tile(rows, M, col_pos) -> DmmaTile:
    rbs = _row_blocks(rows, M, col_pos)
    cost = _tile_cost([len(g) for _, g in rbs])
    assert cost is not None, 'tile exceeds the staging buffer'
    P = cost[1]
    units = [Unit(rb, p, groups[p::P]) for rb, (_, groups) in enumerate(rbs) for p in range(P)]
    order, rounds = _deal([len(u.items) for u in units])
    for k, ui in enumerate(order):
        u = units[ui]
        u.warp, u.rnd = k % NWARP, k // NWARP
        u.slot = u.rb * P + u.part       # parts of a row block sit in adjacent staging slots
    return DmmaTile(rows=np.asarray(rows), parts=P, units=units, round_len=rounds, n_rb=len(rbs))


class DmmaSchedule:
    def __init__(self, tiles: List[DmmaTile], n: int):
        self.tiles, self.n = tiles, n

    # ---- statistics
    @property
    def n_items_real(self):
        return sum(len(u.items) for t in self.tiles for u in t.units)

    @property
    def n_items_padded(self):
        return sum(sum(t.round_len) * NWARP for t in self.tiles)

    @property
    def nnz(self):
        return sum(int(np.count_nonzero(it.A)) for t in self.tiles for u in t.units for it in u.items)

    # ---- numpy restatement of the kernel's executor (w in pivot positions, batch on the leading axes)
    def apply(self, w: np.ndarray) -> np.ndarray:
        w = np.array(w, dtype=float, copy=True)
        for t in self.tiles:
            stage = {}
            for u in t.units:
                acc = np.zeros(w.shape[:-1] + (RB,))
                for it in u.items:
                    acc = acc + w[..., it.pos.astype(int)] @ it.A.T
                stage[u.slot] = acc
            for rb in range(t.n_rb):
                rows = t.rows[rb * RB:(rb + 1) * RB].astype(int)
                tot = sum(stage[rb * t.parts + p] for p in range(t.parts))
                w[..., rows] = tot[..., :len(rows)]
        return w

    # ---- device tables
    def encode(self):
        """Returns dict(tile_hdr int32 (T,4), round_len uint16, items uint32 (n,4,4), vals float64, jobs uint32 (m,4,8)).
        items[i, wg] = [mask, BYTE offset of its first value, row0 | row1 << 16, row2 | row3 << 16] with row_k = w8_row_offset(position k);
        the item stream of a tile is round after round,
        every warp the same number of items per round.  jobs[j, wg] = 8 uint16 row offsets (w8_row_offset; 0xffff: no row), then
        slot0 | parts << 8 | valid << 16 in word 4."""
        tile_hdr, round_len, items, vals, jobs = [], [], [], [], []
        null_item = [0, 0, 0, 0]
        for t in self.tiles:
            item_base, rl_base, job_base = len(items), len(round_len), len(jobs)
            by = {(u.warp, u.rnd): u for u in t.units}
            for r, L in enumerate(t.round_len):
                for i in range(L):
                    row = []
                    for wg in range(NWARP):
                        u = by.get((wg, r))
                        if u is None or i >= len(u.items):
                            row.append(null_item); continue
                        it = u.items[i]
                        flat = it.A.reshape(-1)                      # lane l = row l // 4, k = l % 4
                        mask = 0
                        voff = len(vals)
                        for l in range(32):
                            if flat[l] != 0.0:
                                mask |= 1 << l; vals.append(float(flat[l]))
                        b = [w8_row_offset(int(q)) for q in it.pos]
                        assert max(b) < 65536
                        row.append([mask, voff * 8, b[0] | (b[1] << 16), b[2] | (b[3] << 16)])
                    items.append(row)
                round_len.append(L)
            # staging slot of (warp, round): fixed = wg * n_rounds + r; commit jobs read the slots of their parts
            n_rounds = len(t.round_len)
            slot_of = {u.slot: u.warp * n_rounds + u.rnd for u in t.units}
            rb_jobs = []
            for rb in range(t.n_rb):
                rows = t.rows[rb * RB:(rb + 1) * RB].astype(int)
                pos16 = [w8_row_offset(int(r)) for r in rows] + [0xffff] * (RB - len(rows))     # swizzled byte offsets of the result rows
                slots = [slot_of[rb * t.parts + p] for p in range(t.parts)] + [0] * (4 - t.parts)
                words = [pos16[0] | pos16[1] << 16, pos16[2] | pos16[3] << 16, pos16[4] | pos16[5] << 16, pos16[6] | pos16[7] << 16,
                         slots[0] | slots[1] << 8 | slots[2] << 16 | slots[3] << 24, t.parts, 0, 0]
                rb_jobs.append(words)
            n_jr = -(-len(rb_jobs) // NWARP)
            null_job = [0xffffffff] * 4 + [0, 0, 0, 0]
            for j in range(n_jr):
                jobs.append([rb_jobs[j * NWARP + wg] if j * NWARP + wg < len(rb_jobs) else null_job for wg in range(NWARP)])
            assert n_rounds * NWARP <= MAX_SLOTS * NWARP
            tile_hdr.append([item_base, n_rounds | (n_jr << 8), rl_base, job_base])
        return dict(tile_hdr=np.asarray(tile_hdr, dtype=np.int32).reshape(-1, 4),
                    round_len=np.asarray(round_len, dtype=np.uint16),
                    items=np.asarray(items, dtype=np.uint32).reshape(-1, NWARP, 4),
                    vals=np.asarray(vals + [0.0], dtype=np.float64),
                    jobs=np.asarray(jobs, dtype=np.uint32).reshape(-1, NWARP, 8),
                    max_rounds=max(len(t.round_len) for t in self.tiles))


def build_dmma_schedule(F: LDLFactor, max_group_rows: int = 32) -> DmmaSchedule:
    """Group boundaries by dynamic programming on the tile cost above (same recurrences as schedule.build_schedule, no
    trailing block: every tile is gather-then-write anyway)."""
    level = F.level
    n = len(level)
    bounds = [0] + [k for k in range(1, n) if level[k] != level[k - 1]] + [n]
    nb = len(bounds)
    memo = {}

    def groups(kind, i, j):
        key = (kind, i, j)
        if key not in memo:
            rows, M, cp = (_sched._forward_group if kind == 0 else _sched._backward_group)(F, bounds[i], bounds[j])
            if not len(rows):
                memo[key] = (0.0, [])
            else:
                # a single elimination-tree level may be cut into several tiles (its rows read no other row of the level);
                # a merged group cannot (T^-1 couples its rows), so it has to fit the staging buffer as a whole
                chunk = MAX_SLOTS * RB if j == i + 1 else len(rows)
                parts, tot = [], 0.0
                for r0 in range(0, len(rows), chunk):
                    sub = M[r0:r0 + chunk]
                    c = _tile_cost([len(g) for _, g in _row_blocks(rows[r0:r0 + chunk], sub, cp)])
                    if c is None:
                        tot = float('inf'); break
                    tot += c[0]; parts.append((rows[r0:r0 + chunk], sub, cp))
                memo[key] = (tot, parts)
        return memo[key]
    INF = float('inf')
    best = [[INF] * nb for _ in range(2)]
    prev = [[-1] * nb for _ in range(2)]
    best[0][0] = best[1][0] = 0.0
    for j in range(1, nb):
        for i in range(j - 1, -1, -1):
            if bounds[j] - bounds[i] > max_group_rows and i != j - 1:
                break
            for kind in (0, 1):
                c = best[kind][i] + groups(kind, i, j)[0]
                if c < best[kind][j]:
                    best[kind][j], prev[kind][j] = c, i
    assert best[0][nb - 1] < INF and best[1][nb - 1] < INF, 'a single level exceeds the staging buffer'

    def backtrack(kind):
        out, j = [], nb - 1
        while j > 0:
            out.append((prev[kind][j], j)); j = prev[kind][j]
        return out[::-1]
    tiles = []
    for i, j in backtrack(0):
        tiles += [_make_tile(*g) for g in groups(0, i, j)[1]]
    for i, j in reversed(backtrack(1)):
        tiles += [_make_tile(*g) for g in groups(1, i, j)[1]]
    return DmmaSchedule(tiles, n)


def pack_dmma_blob(S: DmmaSchedule) -> bytes:
    """[CpgDmmaHeader: total_bytes, n_tiles, off_hdr, off_rl, off_items, off_vals, off_jobs, max_rounds] | tile_hdr (int4) | items
    (uint4) | jobs (8 words) | vals (f64) | round_len (u16); every section 16-byte aligned (the blob is staged by TMA bulk copies)."""
    import struct
    E = S.encode()
    pad16 = lambda b: b + b'\0' * ((-len(b)) % 16)
    secs = [pad16(E['tile_hdr'].astype('<i4').tobytes()), pad16(E['items'].astype('<u4').tobytes()), pad16(E['jobs'].astype('<u4').tobytes()),
            pad16(E['vals'].astype('<f8').tobytes()), pad16(E['round_len'].astype('<u2').tobytes())]
    offs, o = [], 32
    for b in secs:
        offs.append(o); o += len(b)
    hdr = struct.pack('<8i', o, len(S.tiles), offs[0], offs[4], offs[1], offs[3], offs[2], int(E['max_rounds']))
    blob = hdr + b''.join(secs)
    assert len(blob) == o and o % 16 == 0
    return blob
