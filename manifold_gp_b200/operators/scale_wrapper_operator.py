"""``ScaleWrapperOperator`` -- manifold_gp/operators/scale_wrapper_operator.py: ``Q * s`` (default) or ``Q / s``
(``inverse_scale``), where Q is the wrapped operator and s the kernel's output scale (``RiemannGP.precision``).

Besides the reference's ``_matmul`` (one inner product + one scaling), the wrapper can run on the solver drivers' caller-owned
buffers (``_mgp_matvec``: no temporaries, CUDA-graph capturable, fused dot product), which is what lets the training loss's
mBCG on ``Noise(Scale(Precision))`` use the same fused iteration as a bare precision operator.  On by default since round 2
(GPU parity against the reference's dense operators: tests/test_gpu_operators.py::test_fused_wrapper_paths_vs_reference_dense);
``MGP_FUSED_WRAPPERS=0`` restores the generic path."""
from __future__ import annotations

import os
from typing import Optional

from torch import Tensor

from .._compat.linear_operator import LinearOperator


def fused_wrappers_enabled() -> bool:
    return os.environ.get("MGP_FUSED_WRAPPERS", "1") != "0"


class ScaleWrapperOperator(LinearOperator):
    def __init__(self, operator: LinearOperator, scale: Tensor, inverse_scale: Optional[bool] = False):
        super().__init__(operator, scale=scale, inverse_scale=inverse_scale)
        self.operator = operator
        self.scale = scale
        self.inverse_scale = inverse_scale

    def _matmul(self, rhs):
        out = self.operator._matmul(rhs.contiguous())                      # :27-28
        return out / self.scale if self.inverse_scale else out * self.scale

    def _size(self):
        return self.operator._size()

    def _transpose_nonbatch(self):
        # the reference passes the bound method instead of calling it (:34, Appendix C.4); every operator on this
        # path is symmetric, so the transpose is the operator built on the transposed inner operator
        return ScaleWrapperOperator(self.operator._transpose_nonbatch(), self.scale, self.inverse_scale)

    # ---- solver-driver interface (see PrecisionMaternOperator._mgp_matvec) ---------------------------------------------------
    def _native(self) -> bool:
        inner = self.operator
        return fused_wrappers_enabled() and hasattr(inner, "_mgp_matvec") and getattr(inner, "_native", lambda: True)()

    def _mgp_structure(self):
        return self.operator._mgp_structure()

    def _mgp_cache_key(self, dtype):
        return ("scale", bool(self.inverse_scale), self.scale.data_ptr(), self.scale._version) + tuple(self.operator._mgp_cache_key(dtype))

    def _coef(self, dtype, outer=None):
        """Device scalar s or 1/s (times an outer wrapper's coefficient), cached per (scale value, outer coefficient)."""
        key = (self.scale.data_ptr(), self.scale._version, dtype, bool(self.inverse_scale),
               None if outer is None else (outer.data_ptr(), outer._version))
        hit = self.__dict__.get("_mgp_coef")
        if hit is None or hit[0] != key:
            import torch
            with torch.no_grad():
                s = self.scale.detach().to(dtype).reshape(-1)[:1]
                c = s.reciprocal() if self.inverse_scale else s.clone()
                if outer is not None:
                    c = c * outer.to(dtype).reshape(-1)[:1]
            hit = (key, c.contiguous(), outer)          # keeps ``outer`` alive: its address is part of the key
            self.__dict__["_mgp_coef"] = hit
        return hit[1]

    def _mgp_matvec(self, x: Tensor, out: Tensor, tmp: Tensor, dot_with=None, dot_out=None, ncols=None, done_flag=None,
                    ep_coef=None, ep_add=None):
        """out[:, :ncols] <- ep_add + ep_coef * s * (Q x) (or / s): the scaling rides in the epilogue of the inner operator's
        last launch (no elementwise pass); the fused dot product is taken with the scaled product."""
        return self.operator._mgp_matvec(x, out, tmp, dot_with=dot_with, dot_out=dot_out, ncols=ncols, done_flag=done_flag,
                                         ep_coef=self._coef(out.dtype, ep_coef), ep_add=ep_add)
