"""Device-resident timing of a family larger than one SM's shared memory (row f3): python tools/time_big.py [name] [B]"""
import json, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
from cvxpygen_b200 import standard
from helpers import family_and_batch

name = sys.argv[1] if len(sys.argv) > 1 else 'random_qp_700_100_700'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
fam, params, _ = family_and_batch(name, B, seed=7)
mod = standard.load(name, device=0)
P = torch.from_numpy(mod.pack_params(params)).cuda()
out = mod.solve_batch_device(P)
torch.cuda.synchronize()
ts = []
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); mod.solve_batch_device(P, out=out); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
it = out.iter.cpu().numpy(); st = out.status.cpu().numpy()
print(json.dumps(dict(name=name, B=B, ms=float(np.median(ts)), inst_per_s=B / (np.median(ts) * 1e-3), mean_iter=float(it.mean()),
                      frac_solved=float((st == 1).mean()), n=mod.dims.n_var, m=mod.dims.n_con)))
