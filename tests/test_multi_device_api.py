"""In-API multi-GPU (VERDICT r1 item 10 / ADVICE r1 medium): one loaded library, one context per device, one host thread per
device inside `cpg_solve_batch_host_multi` -- the product API itself uses the node, not only torchrun.

CPU: the entries are exported, argument errors are codes (not crashes), a second device index does not clobber the first
context (the round-1 bug: `cpg_b200_init(dev1)` re-used device 0's buffers).
GPU: the multi entry on the visible devices returns exactly what the single-device entry returns (any number of devices --
shards are independent); on a multi-GPU box the shards really run on different devices."""
import ctypes as C
import os

import numpy as np
import pytest

from cvxpygen_b200 import standard


def _lib(name):
    return C.CDLL(os.path.join(standard.build(name), 'libcpg_b200.so'))


def test_multi_device_entries_are_exported_and_validate_arguments():
    lib = _lib('mpc_6_3_10')
    for fn in ('cpg_solve_batch_host_multi', 'cpg_b200_use_device', 'cpg_b200_kernel_times'):
        assert hasattr(lib, fn), fn
    assert hasattr(_lib('adp_socp_6_3'), 'cpg_socp_solve_batch_host_multi')
    assert lib.cpg_b200_use_device(C.c_int(3)) == 2          # CPG_B200_ERR_NOT_INIT: that device's context does not exist yet
    assert lib.cpg_b200_use_device(C.c_int(99)) == 3         # CPG_B200_ERR_BAD_ARG
    z = np.zeros(4); zi = np.zeros(4, np.int32)
    p = lambda a, t=C.c_double: a.ctypes.data_as(C.POINTER(t))
    dup = (C.c_int * 2)(0, 0)
    rc = lib.cpg_solve_batch_host_multi(C.c_int(2), dup, C.c_int(4), None, None, None, None, None, None, None, p(z), p(zi, C.c_int),
                                        p(zi, C.c_int), p(z), p(z), None)
    assert rc == 3                                            # the same device twice: one host thread per context
    assert lib.cpg_b200_init(C.c_int(-1)) == 3


@pytest.mark.gpu
def test_multi_entry_equals_single_device_entry():
    import torch
    n = torch.cuda.device_count()
    mod = standard.load('mpc_12_4_10')
    B = 4099
    xi = np.random.default_rng(2).uniform(-1, 1, (B, 12))
    one = mod.solve_batch({'x_init': xi}, return_canonical=True)
    multi = mod.solve_batch_multi({'x_init': xi}, devices=list(range(n)), return_canonical=True)
    assert np.array_equal(multi.sol_x, one.sol_x) and np.array_equal(multi.sol_y, one.sol_y)
    assert np.array_equal(multi.cpg_info.iter, one.cpg_info.iter) and np.array_equal(multi.cpg_info.status, one.cpg_info.status)
    assert np.array_equal(multi.prim, one.prim) and np.array_equal(multi.dual, one.dual)
    ms = standard.load('adp_socp_6_3')
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'socp_adp_socp_6_3.npz'))
    r1 = ms.solve_batch({'f': g['param_f']})
    r2 = ms.solve_batch_multi({'f': g['param_f']}, devices=list(range(n)))
    assert np.array_equal(r1.prim, r2.prim) and np.array_equal(r1.cpg_info.iter, r2.cpg_info.iter)
    if n >= 2:          # both contexts live side by side: interleaved single-device calls on either device agree
        lib = mod.lib
        assert lib.cpg_b200_use_device(C.c_int(1)) == 0
        assert lib.cpg_b200_use_device(C.c_int(0)) == 0
        again = mod.solve_batch({'x_init': xi[:64]}, return_canonical=True)
        assert np.array_equal(again.sol_x, one.sol_x[:64])


@pytest.mark.gpu
def test_multi_entry_with_caller_owned_pinned_outputs():
    """solve_batch_multi(out=...) on numpy views of pinned buffers: same numbers as the allocating call, written in place."""
    import torch
    from cvxpygen_b200 import standard
    mod = standard.load('mpc_6_3_10')
    B = 300
    xi = np.random.default_rng(2).uniform(-1, 1, (B, 6))
    d = mod.dims
    pin = lambda shape, dt=torch.float64: torch.empty(shape, dtype=dt, pin_memory=True).numpy()
    out = dict(prim=pin((B, d.n_prim)), dual=pin((B, d.n_dual)), obj=pin(B), pri=pin(B), dua=pin(B), it=pin(B, torch.int32), st=pin(B, torch.int32))
    devs = list(range(min(2, torch.cuda.device_count())))
    r = mod.solve_batch_multi(xi, devices=devs, out=out)
    ref = mod.solve_batch(xi)
    assert r.prim is out['prim'] and np.array_equal(out['prim'], ref.prim) and np.array_equal(out['dual'], ref.dual)
    assert np.array_equal(out['it'], ref.cpg_info.iter) and np.array_equal(out['st'], ref.cpg_info.status)
    with pytest.raises(ValueError):
        mod.solve_batch_multi(xi, devices=devs, out={**out, 'it': np.zeros(B)})
