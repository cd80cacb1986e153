"""offline/dmma.py: the KKT-solve schedule of the FP64 tensor-core kernel (mma.sync.m8n8k4.f64, instances on N).
(1) `DmmaSchedule.apply` -- the numpy restatement of the kernel's tile executor -- equals K^-1 b for the permuted KKT matrix;
(2) the PACKED tables (what the kernel reads: item descriptors with occupancy masks + compressed coefficients, commit jobs,
    round lengths), decoded here lane by lane exactly as admm_multi_kernel.cuh::dmma_solve indexes them -- fragment layouts of the
    PTX ISA, the w8 swizzle, staging slots, four warps -- give the same result;
(3) every family that takes the tensor-core path keeps within its shared-memory budget (the generated header says so)."""
import struct

import numpy as np
import pytest

from cvxpygen_b200 import families
from cvxpygen_b200.offline.dmma import build_dmma_schedule, pack_dmma_blob, NWARP
from cvxpygen_b200.offline.qp_setup import setup_qp_family

FAMS = [(lambda: families.mpc(6, 3, 10), ['x_init']), (lambda: families.mpc(12, 4, 10), ['x_init']),
        (lambda: families.random_qp(20, 5, 15), ['q', 'b', 'h']), (lambda: families.nonneg_ls(3, 2), ['b'])]


def w8_off(p, c):
    return p * 8 + (c ^ (((p >> 2) & 3) << 1))


def run_tables(blob, w):
    """w: (nk, 8) -> solved in place, following the kernel's indexing (one group: four warps of 32 lanes)."""
    total, n_tiles, off_hdr, off_rl, off_items, off_vals, off_jobs, max_rounds = struct.unpack_from('<8i', blob, 0)
    assert total == len(blob)
    hdr = np.frombuffer(blob, '<i4', n_tiles * 4, off_hdr).reshape(-1, 4)
    rl = np.frombuffer(blob, '<u2', (off_rl and (total - off_rl) // 2), off_rl)
    items = np.frombuffer(blob, '<u4', (off_jobs - off_items) // 4, off_items).reshape(-1, NWARP, 4)
    jobs = np.frombuffer(blob, '<u4', (off_vals - off_jobs) // 4, off_jobs).reshape(-1, NWARP, 8)
    vals = np.frombuffer(blob, '<f8', (off_rl - off_vals) // 8, off_vals)
    nk = w.shape[0]
    w8 = np.zeros((nk + 4) * 8)
    for p in range(nk):
        for c in range(8):
            w8[w8_off(p, c)] = w[p, c]
    lanes = np.arange(32)
    kk, nn = lanes & 3, lanes >> 2
    for t in range(n_tiles):
        item_base, rj, rl_base, job_base = (int(v) for v in hdr[t])
        n_rounds, n_jr = rj & 0xff, rj >> 8
        stage = np.zeros((NWARP * n_rounds, 64))
        for wg in range(NWARP):
            i = item_base
            for r in range(n_rounds):
                L = int(rl[rl_base + r])
                C = np.zeros((8, 8))                                    # result rows x instances
                for _ in range(L):
                    mask, voff, z, wd = (int(v) for v in items[i, wg]); i += 1
                    A = np.zeros(32); B = np.zeros(32)
                    for l in range(32):
                        pw = wd if (kk[l] & 2) else z
                        row = (pw >> (16 * (kk[l] & 1))) & 0xffff           # swizzled byte offset of the operand row
                        B[l] = w8[(row ^ (int(nn[l]) << 3)) >> 3]           # B[k = l % 4][n = l // 4]
                        if (mask >> l) & 1:
                            A[l] = vals[voff // 8 + bin(mask & ((1 << l) - 1)).count('1')]    # A[row = l // 4][k = l % 4]
                    C += A.reshape(8, 4) @ B.reshape(8, 4).T            # B.reshape(8,4)[n][k]
                stage[wg * n_rounds + r] = C.reshape(-1)                # lane l holds C[l // 4][2 (l % 4) + {0, 1}] = flat[2 l + {0, 1}]
        for wg in range(NWARP):
            for j in range(n_jr):
                jb = [int(v) for v in jobs[job_base + j, wg]]
                parts = jb[5]
                acc = sum(stage[(jb[4] >> (8 * q)) & 0xff] for q in range(max(parts, 1)))
                for l in range(32):
                    row = (jb[nn[l] >> 1] >> ((nn[l] & 1) * 16)) & 0xffff      # swizzled byte offset of result row lane // 4
                    if row != 0xffff:
                        o = (row ^ (int(kk[l]) << 4)) >> 3
                        w8[o] = acc[2 * l]; w8[o + 1] = acc[2 * l + 1]
    out = np.zeros_like(w)
    for p in range(nk):
        for c in range(8):
            out[p, c] = w8[w8_off(p, c)]
    return out


@pytest.mark.parametrize('builder,batch', FAMS)
def test_schedule_and_packed_tables_solve_the_kkt_system(builder, batch):
    fam = builder()
    st = setup_qp_family(fam, batch)
    F = st.factor
    nk = st.n + st.m
    S = build_dmma_schedule(F)
    L = np.eye(nk) + F.L
    K = L @ np.diag(F.D) @ L.T
    b = np.random.default_rng(3).standard_normal((8, nk))
    ref = np.linalg.solve(K, b.T).T
    got = S.apply(b)
    scale = np.abs(ref).max()
    assert np.abs(got - ref).max() / scale < 1e-10
    if nk <= 200:                       # the lane-by-lane decoder is slow: small families only
        assert len(st.dmma_blob) > 32
        got2 = run_tables(st.dmma_blob, b.T.copy()).T
        assert np.abs(got2 - ref).max() / scale < 1e-10
    assert S.n_items_padded >= S.n_items_real > 0
    assert all(len(t.round_len) * NWARP <= 64 for t in S.tiles)


def test_tensor_core_path_is_selected_and_fits_shared_memory(tmp_path):
    from cvxpygen_b200 import cpg
    import re
    for name, builder, batch in (('mpc', lambda: families.mpc(12, 4, 10), ['x_init']), ('pf', lambda: families.portfolio_qp(50, 10), ['a', 'w_prev'])):
        d = str(tmp_path / name)
        cpg.generate_code(builder(), code_dir=d, batch_params=batch, wrapper=False, solver_opts={'dmma': True})
        h = dict(re.findall(r'#define (CPG_FAM_\w+) (\d+)', open(f'{d}/c/include/cpg_family.h').read()))
        h = {k: int(v) for k, v in h.items()}
        assert name != 'mpc' or h['CPG_FAM_DMMA'] == 1          # opted in: the headline family fits the tensor-core kernel
        if not h['CPG_FAM_DMMA']:                                # too large for two groups next to its tables: straight-line kernel
            continue
        assert 2 <= h['CPG_FAM_DM_GROUPS'] <= 3
        smem = h['CPG_FAM_CBLOB_BYTES_PAD'] + h['CPG_FAM_DBLOB_BYTES_PAD'] + h['CPG_FAM_DM_GROUPS'] * (
            (h['CPG_FAM_DM_W8'] + h['CPG_FAM_DM_STAGE']) * 8 + 4 * h['CPG_FAM_DM_BV'] * 8 + 8) + 16
        assert smem <= 232448 - 1024
    # a family with per-instance matrices has no shared factor: the tensor-core path stays off even when asked for; so does the default
    d = str(tmp_path / 'ltv')
    cpg.generate_code(families.mpc_ltv(4, 2, 5), code_dir=d, batch_params=['A', 'B', 'qdiag', 'rdiag', 'x_init'], wrapper=False, solver_opts={'dmma': True})
    assert '#define CPG_FAM_DMMA 0' in open(f'{d}/c/include/cpg_family.h').read()
    d = str(tmp_path / 'default')
    cpg.generate_code(families.mpc(4, 2, 6), code_dir=d, batch_params=['x_init'], wrapper=False)
    assert '#define CPG_FAM_DMMA 0' in open(f'{d}/c/include/cpg_family.h').read()
