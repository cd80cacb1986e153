"""Config 3 at full size against the compiled ECOS: which instances differ in exit flag / iteration count, and by how much."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from cvxpygen_b200 import standard
wl = bench.WORKLOADS['portfolio_socp']
B = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
P = wl.host_params(B, 1)
mod = standard.load(wl.family)
res = mod.solve_batch(P, return_canonical=True)
fam, ora = wl.reference(P, os.cpu_count())
rel = lambda a, b: np.linalg.norm(a - b, axis=1) / np.maximum(np.linalg.norm(b, axis=1), 1e-12)
prim_ref = np.concatenate([ora['x'][:, v.indices] for v in fam.variables], axis=1)
dual_ref = np.concatenate([ora[d.vec][:, d.indices] for d in fam.duals], axis=1)
ep, ed = rel(res.prim, prim_ref), rel(res.dual, dual_ref)
ex, ey, es, ez = rel(res.sol_x, ora['x']), rel(res.sol_y, ora['y']), rel(res.sol_s, ora['s']), rel(res.sol_z, ora['z'])
st, ef = res.cpg_info.status, ora['exitflag']
it, ir = res.cpg_info.iter.astype(np.int64), ora['iter']
bad = np.nonzero(st != ef)[0]
out = dict(B=B, status_mismatch=int(len(bad)), pairs=[(int(i), int(st[i]), int(ef[i]), int(it[i]), int(ir[i]), float(ep[i]), float(ed[i])) for i in bad[:20]],
           iter_diff_hist={int(k): int(v) for k, v in zip(*np.unique(it - ir, return_counts=True))},
           max_rel=dict(prim=float(ep.max()), dual=float(ed.max()), x=float(ex.max()), y=float(ey.max()), s=float(es.max()), z=float(ez.max())),
           max_rel_where_status_equal=dict(prim=float(ep[st == ef].max()), dual=float(ed[st == ef].max())),
           quantiles_z={q: float(np.quantile(ez, q)) for q in (0.5, 0.99, 0.999, 0.9999)},
           quantiles_dual={q: float(np.quantile(ed, q)) for q in (0.5, 0.99, 0.999, 0.9999)},
           ref_flags={int(k): int(v) for k, v in zip(*np.unique(ef, return_counts=True))},
           our_flags={int(k): int(v) for k, v in zip(*np.unique(st, return_counts=True))})
print(json.dumps(out))
