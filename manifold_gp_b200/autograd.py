"""Autograd glue for the CUDA kernels.

The reference gets its gradients by autograd tracing through the torch ops inside ``_matmul`` and the cached value
properties (graph_laplacian_operator.py:52-124; tested by test_grad / test_ml, test/_test_functions.py:59-104).
Here the same derivatives are hand-written kernels:

* d/d(rhs)              -> the SpMM itself (the operator is symmetric up to the pre/post scalings),
* d/d(a, diag, shift)   -> ``mgp_lap_sddmm``  (what autograd through ``torch_sparse.spmm`` yields for ``value``),
* d/d(eps)              -> ``mgp_lap_values_grad`` (forward-mode tangents of the value build, fused reduction).
"""
from __future__ import annotations

import torch

from . import _lib, graph
from ._lib import c_int32, c_int64, ptr, stream


class _LapValuesFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, eps, st, d2csr, self_loops):
        deg_un, deg, diag, a = graph.lap_values(st, d2csr, eps, self_loops)
        ctx.st, ctx.d2, ctx.self_loops = st, d2csr, self_loops
        ctx.eps_shape, ctx.eps_dtype = eps.shape, eps.dtype
        ctx.save_for_backward(eps.detach(), deg_un, deg, diag, a)
        return deg_un, deg, diag, a

    @staticmethod
    def backward(ctx, g_dt, g_dg, g_diag, g_a):
        eps, deg_un, deg, diag, a = ctx.saved_tensors
        st = ctx.st
        dt = a.dtype
        dev = a.device
        if g_diag is None:
            g_diag = torch.zeros_like(diag)
        if g_a is None:
            g_a = torch.zeros_like(a)
        eps_t = graph._device_scalar(eps, dt, dev)
        out = torch.zeros(1, dtype=dt, device=dev)
        nb = _lib.query("mgp_lap_values_grad_ws_bytes", c_int64(st.n))
        ws = torch.zeros(nb, dtype=torch.uint8, device=dev)
        _lib.call("mgp_lap_values_grad_" + _lib.suffix(dt), ptr(st.rowptr), ptr(st.col), ptr(ctx.d2), c_int64(st.n),
                  ptr(eps_t), c_int32(1 if ctx.self_loops else 0), ptr(deg_un), ptr(deg), ptr(diag), ptr(a),
                  ptr(None if g_dt is None else g_dt.contiguous()), ptr(None if g_dg is None else g_dg.contiguous()),
                  ptr(g_diag.contiguous()), ptr(g_a.contiguous()), ptr(out), ptr(ws), stream())
        return out.reshape(ctx.eps_shape).to(ctx.eps_dtype), None, None, None


def lap_values_autograd(st, d2csr, eps, self_loops):
    return _LapValuesFn.apply(eps, st, d2csr, self_loops)


class _LapSpmmFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, diag, x, shift, pre, post, st, x_external, y_external):
        y = graph.lap_spmm(st, a, diag, x, shift=shift, pre=pre, post=post, x_external=x_external, y_external=y_external)
        ctx.st, ctx.xe, ctx.ye = st, x_external, y_external
        ctx.save_for_backward(a, diag, x, shift, pre, post, y if (post is not None and post.requires_grad) else None)
        return y

    @staticmethod
    def backward(ctx, gy):
        a, diag, x, shift, pre, post, y = ctx.saved_tensors
        st = ctx.st
        gy = gy.contiguous()
        need = ctx.needs_input_grad
        g_a = g_diag = g_x = g_shift = g_pre = g_post = None
        # internal-order copies for the pieces that index per-node arrays
        gy_i = st.to_internal(gy) if ctx.ye else gy
        x_i = st.to_internal(x) if ctx.xe else x
        if need[2] or need[4]:
            # dL/dZ with Z = pre .* X:  M (post .* gy)   (M symmetric); result in X's row order
            dz = graph.lap_spmm(st, a, diag, gy, shift=shift, pre=post, post=None, x_external=ctx.ye, y_external=False)
            if need[4] and pre is not None:
                g_pre = (dz * x_i).sum(1)
            if need[2]:
                g_x = dz if pre is None else dz * pre.view(-1, 1)
                if ctx.xe:
                    g_x = st.to_external(g_x)
        if need[0] or need[1] or need[3]:
            g_a, g_diag = graph.lap_sddmm(st, gy_i, x_i, pre=pre, post=post)
            if need[3] and shift is not None:
                g_shift = g_diag.sum().reshape(shift.shape)
        if need[5] and post is not None:
            y_i = st.to_internal(y) if ctx.ye else y
            g_post = (gy_i * y_i).sum(1) / post
        return g_a, g_diag, g_x, g_shift, g_pre, g_post, None, None, None


def lap_spmm_apply(st, a, diag, x, shift, pre, post, x_external=False, y_external=False):
    """Differentiable ``Y = post .* ((diag + shift) .* (pre .* X) - A (pre .* X))`` (plain kernel call when no
    input requires grad).  ``x_external`` / ``y_external``: X / Y in the caller's row order (see graph.lap_spmm)."""
    tensors = (a, diag, x, shift, pre, post)
    if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors):
        return _LapSpmmFn.apply(a, diag, x, shift, pre, post, st, x_external, y_external)
    return graph.lap_spmm(st, a, diag, x, shift=shift, pre=pre, post=post, x_external=x_external, y_external=y_external)
