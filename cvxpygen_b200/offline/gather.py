"""Gather plans: the table format behind every sparse kernel phase of the IPM-CUDA backend.

A *phase* is a set of independent rows  out[target] (op)= sum_e value(entry_e)  that all threads of a CTA work on between
two barriers (one level of the numeric LDL', one level of a triangular solve, one product with the KKT matrix).  The
reference walks compressed columns one after the other on one core (ecos/external/ldl/src/ldl.c:266-360, :362-507,
ecos/src/spla.c:22-76); a CTA needs the opposite: every target owned by exactly one writer (no atomics, fixed summation
order -> bitwise reproducible), index tables read coalesced and conflict-free, no data-dependent control flow.

Layout produced here, per phase:
  * rows are cut into CELLS of at most q entries; a row of c entries takes g = 2^ceil(log2(ceil(c/q))) cells (<= 32) that
    sit on g adjacent lanes of one warp (aligned to g) and are summed with log2(g) butterfly shuffles;
  * cells are sorted by (g, length) descending and dealt T at a time into ROUNDS; inside a round every cell is padded
    with null entries to the round's length, so the entry loop has a uniform trip count and entry j of thread t sits
    at  base + j*T + t  (coalesced, bank-conflict free);
  * one descriptor per (round, thread): target | log2(g) << GSHIFT | valid << VSHIFT | flag << FSHIFT.
q is chosen per phase by a small cost model (rounds x (fixed + length + shuffle steps)).
"""
from dataclasses import dataclass, field
from typing import List, Sequence, Tuple

import numpy as np


@dataclass
class Phase:
    round_lo: int                 # first round of the phase
    round_hi: int


@dataclass
class GatherPlan:
    T: int
    nfields: int
    tbits: int                    # bits of the target field in a descriptor
    null_entry: Tuple[int, ...]
    desc: List[int] = field(default_factory=list)            # (n_rounds * T)
    entries: List[Tuple[int, ...]] = field(default_factory=list)   # per (round, warp): len x 32 entries, lane-interleaved
    wr_base: List[int] = field(default_factory=list)         # per (round, warp): entry offset, trip count, shuffle steps
    wr_len: List[int] = field(default_factory=list)
    wr_shuf: List[int] = field(default_factory=list)
    phases: List[Phase] = field(default_factory=list)

    @property
    def nwarp(self):
        return self.T // 32

    @property
    def gshift(self):
        return self.tbits

    @property
    def vshift(self):
        return self.tbits + 3

    @property
    def fshift(self):
        return self.tbits + 4

    @property
    def n_rounds(self):
        return len(self.wr_len) // self.nwarp

    def desc_array(self):
        dt = np.uint16 if self.tbits + 5 <= 16 else np.uint32
        return np.asarray(self.desc, dtype=np.int64).astype(dt)

    def entry_array(self):
        """(n_entries, nfields) uint16"""
        return np.asarray(self.entries, dtype=np.int64).reshape(-1, self.nfields).astype(np.uint16)


def _next_pow2(v: int) -> int:
    r = 1
    while r < v:
        r *= 2
    return r


def _cells_for(rows, q):
    cells = []      # (g, length, row index, part)
    for ri, (_, _, ent) in enumerate(rows):
        c = len(ent)
        g = min(32, _next_pow2(-(-c // q))) if c else 1
        ln = -(-c // g) if c else 0
        for part in range(g):
            cells.append((g, ln, ri, part))
    cells.sort(key=lambda x: (-x[0], -x[1], x[2], x[3]))
    return cells


def _deal(cells, nwarp):
    """Chunks of 32 sorted cells -> (round, warp), snaking so that every warp gets its share of the long cells."""
    out = []
    for c in range(0, len(cells), 32):
        i = c // 32
        r, pos = divmod(i, nwarp)
        out.append((r, pos if r % 2 == 0 else nwarp - 1 - pos, cells[c:c + 32]))
    return out


import os as _os
_C_FIXED = float(_os.environ.get('CPG_GATHER_C_FIXED', 10.0))      # (environment overrides: A/B sweeps of the cell-size model only)
_C_SHUF = float(_os.environ.get('CPG_GATHER_C_SHUF', 3.0))


def _cost(cells, nwarp, c_fixed=_C_FIXED, c_entry=1.0, c_shuf=_C_SHUF):
    per_warp = [0.0] * nwarp
    for r, w, chunk in _deal(cells, nwarp):
        per_warp[w] += c_fixed + c_entry * max(c[1] for c in chunk) + c_shuf * (max(c[0] for c in chunk).bit_length() - 1)
    return max(per_warp)


# Deal a lane's entries to the steps so that half-warps hit distinct bank pairs (_order_steps).  OFF: with cells this short there is
# little freedom -- the bank model counts 3 404 -> 2 566 extra passes per iteration for portfolio_socp_100_10 (-25 % of the excess,
# ~ -8 % of the operand wavefronts of a kernel whose shared-memory pipe is 31 % busy) -- not worth a different summation order.
CONFLICT_AWARE = False


def _order_steps(lists, ln, nfields, null_entry):
    """The ORDER in which a lane walks the entries of its cell is free (a sum).  An 8-byte shared-memory load of a warp is served
    in two passes, one per half-warp, iff the distinct addresses of a half fall into distinct bank pairs (index mod 16); every
    extra address in a pair costs a pass (ncu, round 2: 22 % of ipm_kernel's shared wavefronts were such conflicts, half of
    the operand gathers').  Greedy, step by step and half by half: the lanes with the most entries left choose first, each takes
    the entry that adds the fewest conflicts over all operand fields; a lane with slack may wait (its step is a null entry).
    Returns ln lists of 32 entries (None = null)."""
    rem = [list(l) for l in lists]
    out = []
    nf = nfields
    for j in range(ln):
        left = ln - j
        step = [None] * 32
        for h in (0, 16):
            used = [dict() for _ in range(nf)]       # per field: bank pair -> set of addresses
            lanes = sorted(range(h, h + 16), key=lambda l: -len(rem[l]))
            for l in lanes:
                if not rem[l]:
                    continue
                best_i, best_c = -1, None
                for i, e in enumerate(rem[l]):
                    c = 0
                    for f in range(nf):
                        a = e[f]; cls = used[f].get(a & 15)
                        if cls and a not in cls:
                            c += 1
                    if best_c is None or c < best_c:
                        best_i, best_c = i, c
                        if c == 0:
                            break
                if best_c and len(rem[l]) < left:
                    continue                       # would conflict and the lane has slack: wait for a later step
                e = rem[l].pop(best_i)
                for f in range(nf):
                    used[f].setdefault(e[f] & 15, set()).add(e[f])
                step[l] = e
        out.append(step)
    assert not any(rem)
    return out


def plan_conflicts(plan: GatherPlan) -> int:
    """Extra shared-memory passes of the plan's operand loads under the bank model above (diagnostic / tests)."""
    E = plan.entry_array().astype(np.int64).reshape(-1, 32, plan.nfields)
    extra = 0
    for step in E:
        for h in (0, 16):
            for f in range(plan.nfields):
                a = np.unique(step[h:h + 16, f])
                extra += int(np.bincount(a & 15, minlength=16).max()) - 1
    return extra


def add_phase(plan: GatherPlan, rows: Sequence[Tuple[int, int, Sequence[Tuple[int, ...]]]], use_warps: int = 0) -> None:
    """rows: (target, flag, entries).  Appends one phase (possibly of zero rounds) to the plan.  use_warps > 0 deals the
    cells to the first use_warps warps only (the others are busy with something else during this phase)."""
    T, NW = plan.T, plan.nwarp
    NU = use_warps if 0 < use_warps < NW else NW
    lo = plan.n_rounds
    rows = [(int(t), int(f), [tuple(int(v) for v in e) for e in ent]) for t, f, ent in rows]
    if not rows:
        plan.phases.append(Phase(lo, lo))
        return
    cmax = max(len(r[2]) for r in rows)
    qmin = max(1, -(-cmax // 32))
    best = None
    for q in range(qmin, max(qmin, min(cmax, 64)) + 1):
        cells = _cells_for(rows, q)
        cost = _cost(cells, NU)
        if best is None or cost < best[0]:
            best = (cost, q, cells)
    cells = best[2]
    assert all(t < (1 << plan.tbits) for t, _, _ in rows)
    dealt = _deal(cells, NU)
    n_rounds = dealt[-1][0] + 1
    desc = [0] * (n_rounds * T)
    wr = {(r, w): chunk for r, w, chunk in dealt}
    for r in range(n_rounds):
        for w in range(NW):
            chunk = wr.get((r, w), [])
            ln = max([c[1] for c in chunk], default=0)
            shuf = max([c[0] for c in chunk], default=1).bit_length() - 1
            plan.wr_base.append(len(plan.entries)); plan.wr_len.append(ln); plan.wr_shuf.append(shuf)
            ent = [plan.null_entry] * (ln * 32)
            lists = [[] for _ in range(32)]
            for lane, (g, clen, ri, part) in enumerate(chunk):
                assert lane % g == part                   # the group is aligned to its size inside the warp
                tgt, flag, e = rows[ri]
                lists[lane] = list(e[part * clen:(part + 1) * clen])
                desc[r * T + w * 32 + lane] = tgt | ((g.bit_length() - 1) << plan.gshift) | (1 << plan.vshift) | (flag << plan.fshift)
            for j, step in enumerate(_order_steps(lists, ln, plan.nfields, plan.null_entry) if CONFLICT_AWARE else
                                     [[lst[j] if j < len(lst) else None for lst in lists] for j in range(ln)]):
                for lane, v in enumerate(step):
                    if v is not None:
                        ent[j * 32 + lane] = v
            plan.entries += ent
    plan.desc += desc
    plan.phases.append(Phase(lo, lo + n_rounds))


def run_phase(plan: GatherPlan, phase: int, value, commit) -> None:
    """Pure-Python executor with the kernel's summation order: value(entry tuple) -> float, commit(target, flag, acc)."""
    T, NW = plan.T, plan.nwarp
    ph = plan.phases[phase]
    for r in range(ph.round_lo, ph.round_hi):
        for w in range(NW):
            i = r * NW + w
            base, ln, shuf = plan.wr_base[i], plan.wr_len[i], plan.wr_shuf[i]
            acc = np.zeros(32)
            for j in range(ln):
                for lane in range(32):
                    acc[lane] += value(plan.entries[base + j * 32 + lane])
            d = plan.desc[r * T + w * 32:r * T + w * 32 + 32]
            g = np.array([1 << ((x >> plan.gshift) & 7) for x in d])
            o = 1
            for _ in range(shuf):
                other = acc[np.arange(32) ^ o]
                acc = np.where(o < g, acc + other, acc)
                o *= 2
            for lane in range(32):
                if (d[lane] >> plan.vshift) & 1 and lane % g[lane] == 0:
                    commit(d[lane] & ((1 << plan.tbits) - 1), (d[lane] >> plan.fshift) & 1, acc[lane])
