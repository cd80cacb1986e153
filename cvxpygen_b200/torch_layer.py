"""Batched differentiable QP layer on top of a generated ADMM-CUDA library.

The reference exposes `forward(params, context)` / `backward(dvars, context)` for cvxpylayers' `custom_method`
(cvxpygen/templates/cpg_solver.py.jinja2:176-212), one instance per call.  This is the batched equivalent as a
torch.autograd.Function: forward = cpg_solve_batch_device, backward = cpg_gradient_batch_device (or its _mat variant for families with
per-instance matrix parameters, whose gradients include dP / dA folded through the P / A maps); everything stays on
the GPU (the canonical dual of the forward pass is kept for the backward pass, like `gradient_dual` in the reference)."""
import torch


class _BatchedQP(torch.autograd.Function):
    @staticmethod
    def forward(ctx, params, module):
        p = params.detach().contiguous()
        out = module.solve_batch_device(p, return_canonical=True)
        ctx.module = module
        if module.has_matrix_params:          # dP / dA need the parameter rows and the primal solution as well
            ctx.save_for_backward(out.sol_y, out.sol_x, p)
        else:
            ctx.save_for_backward(out.sol_y)
        ctx.status = out.status
        return out.prim

    @staticmethod
    def backward(ctx, dprim):
        if ctx.module.has_matrix_params:
            sol_y, sol_x, p = ctx.saved_tensors
            return ctx.module.gradient_batch_device_mat(p, sol_x, sol_y, dprim.contiguous()), None
        (sol_y,) = ctx.saved_tensors
        return ctx.module.gradient_batch_device(sol_y, dprim.contiguous()), None


class BatchedQPLayer(torch.nn.Module):
    """params (B, n_param) float64 CUDA tensor of the batched user parameters -> (B, n_prim) user variables."""

    def __init__(self, module):
        super().__init__()
        self.module = module.init()

    def forward(self, params):
        return _BatchedQP.apply(params, self.module)
