"""One warm-up + one device-resident solve of a standard family (target of ncu): python tools/prof_family.py <name> [B]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
from cvxpygen_b200 import standard

name = sys.argv[1]; B = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
mod = standard.load(name, device=0)
if name in standard.MATPAR_NAMES:
    from helpers import ltv_batch
    params = ltv_batch(standard.STANDARD[name][0](), B, seed=31)
elif name in standard.SOCP_NAMES:
    import bench
    params = bench.WORKLOADS['portfolio_socp'].host_params(B, 1)
else:
    from helpers import family_and_batch
    params = family_and_batch(name, B, seed=1)[1]
P = torch.from_numpy(np.ascontiguousarray(mod.pack_params(params) if isinstance(params, dict) else params)).cuda()
out = mod.solve_batch_device(P)
torch.cuda.synchronize()
out = mod.solve_batch_device(P, out=out)
torch.cuda.synchronize()
print(name, B, float(out.iter.float().mean()), mod.kernel_times() if hasattr(mod, 'kernel_times') else '')
