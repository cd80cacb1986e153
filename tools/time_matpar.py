"""Device-resident timing of the matrix-parameter kernel (row f2): python tools/time_matpar.py [name] [B]"""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np, torch
from cvxpygen_b200 import standard
from helpers import ltv_batch

name = sys.argv[1] if len(sys.argv) > 1 else 'mpc_ltv_12_4_10'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
fam = standard.STANDARD[name][0]()
mod = standard.load(name, device=0)
P = torch.from_numpy(mod.pack_params(ltv_batch(fam, B, seed=31))).cuda()
out = mod.solve_batch_device(P)
torch.cuda.synchronize()
ts = []
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); mod.solve_batch_device(P, out=out); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
it = out.iter.cpu().numpy(); st = out.status.cpu().numpy()
print(json.dumps(dict(name=name, B=B, ms=float(np.median(ts)), inst_per_s=B / (np.median(ts) * 1e-3), mean_iter=float(it.mean()),
                      frac_solved=float((st == 1).mean()))))
