#!/usr/bin/env python
"""Diagnostic: instances of the config-3 batch whose exit flag is not 0 on the GPU; saves them with the reference's answer."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), 'tests'))
import torch
from cvxpygen_b200 import standard, families
from oracle import ref_ecos
fam = families.portfolio_socp()
B = 50000
rng = np.random.default_rng(11)
a = rng.standard_normal((B, 100)); wp = np.abs(1 / 100 + 0.01 * rng.standard_normal((B, 100)))
m = standard.load('portfolio_socp_100_10')
P = torch.from_numpy(np.ascontiguousarray(np.c_[a, wp])).cuda()
out = m.solve_batch_device(P, return_canonical=True)
torch.cuda.synchronize()
st = out.status.cpu().numpy(); it = out.iter.cpu().numpy()
odd = np.nonzero(st != 0)[0]
print('odd', odd, st[odd], it[odd], 'pres', out.pri_res.cpu().numpy()[odd], 'dres', out.dua_res.cpu().numpy()[odd])
c0, b0 = fam.canon_data('c'), fam.canon_data('b')
r = ref_ecos.RefECOS(c0, fam.canon_matrix('A'), b0, fam.canon_matrix('G'), fam.canon_data('h'), 601, [12, 102])
n = len(odd)
Cb = np.tile(c0, (n, 1)); Cb[:, :100] = -a[odd]
Bb = np.tile(b0, (n, 1)); Bb[:, 11:111] = -wp[odd]
ref = r.solve_batch(c=Cb, b=Bb)
print('ref flags', ref['exitflag'], 'iters', ref['iter'], 'pres', ref['pres'], 'dres', ref['dres'])
np.savez('gpurun_out/diag_socp2.npz', odd=odd, a=a[odd], wp=wp[odd], st=st[odd], it=it[odd], x=out.sol_x.cpu().numpy()[odd], refx=ref['x'], refit=ref['iter'], refflag=ref['exitflag'])
