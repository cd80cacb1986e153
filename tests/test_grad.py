"""Backward pass (gradient=True).
CPU: the numpy restatement (oracle/grad_numpy.py) against the reference's own generated C (golden vectors from
oracle/build_grad_ref.py) and against finite differences of the solution map.
GPU: the sm_100a kernel (through the C ABI) against both."""
import os

import numpy as np
import pytest

from cvxpygen_b200 import standard
from helpers import GOLDEN, family_and_batch, oracle_solve, canon_batches
from oracle.grad_numpy import qp_backward, param_gradient

GRAD_FAMS = ['mpc_12_4_10', 'mpc_6_3_10', 'nonneg_LS_3_2']
TOL = 1e-5


def relmax(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


@pytest.mark.parametrize('name', GRAD_FAMS)
def test_numpy_backward_matches_reference_generated_c(name):
    g = np.load(os.path.join(GOLDEN, f'grad_{name}.npz'))
    fam = standard.STANDARD[name][0]()
    n = fam.n_var
    prim_idx = np.concatenate([v.indices for v in fam.variables])
    dx = np.zeros((g['dprim'].shape[0], n)); dx[:, prim_idx] = g['dprim']
    dq, dl, du, _ = qp_backward(fam.canon_matrix('P'), fam.canon_matrix('A'), g['sol_x'], g['sol_y'], dx)
    assert relmax(dq, g['dq']) < 1e-9
    assert relmax(dl + du, g['dl'] + g['du']) < 1e-9
    names = standard.STANDARD[name][1]
    assert relmax(param_gradient(fam, dq, dl, du, names), param_gradient(fam, g['dq'], g['dl'], g['du'], names)) < 1e-9


def test_backward_matches_finite_differences():
    """d/dtheta of c'x*(theta) for the MPC family (theta = x_init), strict complementarity, tight forward solves."""
    name = 'mpc_6_3_10'
    fam, params, (q, l, u) = family_and_batch(name, 4, seed=9)
    kw = dict(eps_abs=1e-11, eps_rel=1e-11, max_iter=200000, adaptive_rho=1)
    sol = oracle_solve(fam, q, l, u, prefer_ref=True, **kw)
    cvec = np.random.default_rng(1).standard_normal(fam.n_var)
    dx = np.tile(cvec, (4, 1))
    dq, dl, du, _ = qp_backward(fam.canon_matrix('P'), fam.canon_matrix('A'), sol['x'], sol['y'], dx)
    g = param_gradient(fam, dq, dl, du, ['x_init'])
    h = 1e-5
    fd = np.zeros_like(g)
    for k in range(g.shape[1]):
        for sgn in (+1, -1):
            p2 = {'x_init': params['x_init'].copy()}; p2['x_init'][:, k] += sgn * h
            q2, l2, u2 = canon_batches(fam, p2, 4)
            s2 = oracle_solve(fam, q2, l2, u2, prefer_ref=True, **kw)
            fd[:, k] += sgn * (s2['x'] @ cvec) / (2 * h)
    assert relmax(g, fd) < 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize('name', GRAD_FAMS)
def test_gpu_backward_matches_reference_golden(name):
    g = np.load(os.path.join(GOLDEN, f'grad_{name}.npz'))
    fam = standard.STANDARD[name][0]()
    mod = standard.load(name)
    res, dq, dl, du = mod.gradient_batch(g['sol_y'], g['dprim'], return_canonical=True)
    assert relmax(dq, g['dq']) < TOL
    assert relmax(dl + du, g['dl'] + g['du']) < TOL
    names = standard.STANDARD[name][1]
    ref = param_gradient(fam, g['dq'], g['dl'], g['du'], names)
    got = np.concatenate([res[nm] for nm in names], axis=1)
    assert relmax(got, ref) < TOL


@pytest.mark.gpu
def test_gpu_forward_backward_pipeline_and_torch_layer():
    """Config 4: solve, then differentiate, on device tensors; also through the autograd wrapper."""
    import torch
    from cvxpygen_b200.torch_layer import BatchedQPLayer
    name, B = 'mpc_12_4_10', 512
    fam, params, (q, l, u) = family_and_batch(name, B, seed=21)
    mod = standard.load(name)
    layer = BatchedQPLayer(mod)
    th = torch.tensor(params['x_init'], dtype=torch.float64, device='cuda', requires_grad=True)
    prim = layer(th)                                        # (B, n_prim) user variables U then X
    wgt = torch.randn(prim.shape, dtype=torch.float64, device='cuda', generator=torch.Generator('cuda').manual_seed(0))
    (prim * wgt).sum().backward()
    sol = oracle_solve(fam, q, l, u)
    prim_idx = np.concatenate([v.indices for v in fam.variables])
    assert np.allclose(prim.detach().cpu().numpy(), sol['x'][:, prim_idx], rtol=1e-6, atol=1e-9)
    dx = np.zeros((B, fam.n_var)); dx[:, prim_idx] = wgt.cpu().numpy()
    dq, dl, du, _ = qp_backward(fam.canon_matrix('P'), fam.canon_matrix('A'), sol['x'][:64], sol['y'][:64], dx[:64])
    ref = param_gradient(fam, dq, dl, du, ['x_init'])
    assert relmax(th.grad[:64].cpu().numpy(), ref) < TOL


@pytest.mark.gpu
def test_gpu_backward_is_bit_reproducible():
    """The per-instance numeric factorisation runs without atomics (coloured rounds, csrc/admm_kernel.cuh:tail_factor): two backward
    passes over the same 20 000 solutions return the same bits."""
    name, B = 'mpc_12_4_10', 20000
    fam, params, _ = family_and_batch(name, B, seed=4)
    mod = standard.load(name)
    res = mod.solve_batch(params, return_canonical=True)
    dprim = np.random.default_rng(0).standard_normal((B, mod.dims.n_prim))
    a = mod.gradient_batch(res.sol_y, dprim, return_canonical=True)
    b = mod.gradient_batch(res.sol_y, dprim, return_canonical=True)
    for u, v in zip(a[1:], b[1:]):
        assert np.array_equal(u, v)
    assert all(np.array_equal(a[0][k], b[0][k]) for k in a[0])
