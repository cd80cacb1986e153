"""One device-resident solve of a tools/_variants/<name> library (for ncu): python tools/prof_variant.py <variant> [B]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from cvxpygen_b200 import runtime
v = sys.argv[1]; B = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
mod = runtime.Module(os.path.join(ROOT, 'tools', '_variants', v)).init()
P = torch.from_numpy(np.random.default_rng(1).uniform(-1, 1, (B, mod.dims.n_param))).cuda()
out = mod.solve_batch_device(P)
torch.cuda.synchronize()
print(v, float(out.iter.float().mean()), mod.kernel_times())
