"""``NoiseWrapperOperator`` -- manifold_gp/operators/noise_wrapper_operator.py: the 3-term Neumann series of
``(Q^-1 + s I)^-1``, i.e. ``Q - s Q^2 + s^2 Q^3`` evaluated as three nested products with the wrapped operator Q
(meaningful while ``s * |Q| < 1``; the reference's dense twin is ``test/_dense_operators.py:56-57``).

Like ``ScaleWrapperOperator`` it also offers the solver drivers' buffer interface (``_mgp_matvec``; ``MGP_FUSED_WRAPPERS=0``
switches it off): three inner fused products and two in-place axpy passes on two cached scratch blocks."""
from __future__ import annotations

from torch import Tensor

from .._compat.linear_operator import LinearOperator
from .scale_wrapper_operator import fused_wrappers_enabled


class NoiseWrapperOperator(LinearOperator):
    def __init__(self, operator: LinearOperator, noise: Tensor):
        super().__init__(operator, noise=noise)
        self.operator = operator
        self.noise = noise
        self._mgp_scratch = None

    def _matmul(self, rhs):
        rhs = rhs.contiguous()
        op = self.operator._matmul                                           # :21-22
        return op(rhs - self.noise * op(rhs - self.noise * op(rhs)))

    def _size(self):
        return self.operator._size()

    def _transpose_nonbatch(self):
        return NoiseWrapperOperator(self.operator._transpose_nonbatch(), self.noise)   # see ScaleWrapperOperator

    # ---- solver-driver interface ---------------------------------------------------------------------------------------------
    def _native(self) -> bool:
        inner = self.operator
        return fused_wrappers_enabled() and hasattr(inner, "_mgp_matvec") and getattr(inner, "_native", lambda: True)()

    def _mgp_structure(self):
        return self.operator._mgp_structure()

    def _mgp_cache_key(self, dtype):
        return ("noise", self.noise.data_ptr(), self.noise._version) + tuple(self.operator._mgp_cache_key(dtype))

    def _scratch(self, like: Tensor):
        key = (tuple(like.shape), like.dtype, like.device)
        if self._mgp_scratch is None or self._mgp_scratch[0] != key:
            import torch
            self._mgp_scratch = (key, torch.zeros_like(like), torch.zeros_like(like))
        return self._mgp_scratch[1], self._mgp_scratch[2]

    def _neg_noise(self, dtype):
        key = (self.noise.data_ptr(), self.noise._version, dtype)
        hit = self.__dict__.get("_mgp_negs")
        if hit is None or hit[0] != key:
            import torch
            with torch.no_grad():
                hit = (key, (-self.noise.detach().to(dtype).reshape(-1)[:1]).contiguous())
            self.__dict__["_mgp_negs"] = hit
        return hit[1]

    def _mgp_matvec(self, x: Tensor, out: Tensor, tmp: Tensor, dot_with=None, dot_out=None, ncols=None, done_flag=None,
                    ep_coef=None, ep_add=None):
        """out <- Q (x - s Q (x - s Q x)) on caller-owned [n, ld] buffers: three inner fused products; the two ``x - s Q(.)``
        combinations ride in the epilogue of the inner launches (ep_coef = -s, ep_add = x), the dot product comes out of the
        last one."""
        inner = self.operator._mgp_matvec
        u, w = self._scratch(x)
        ns = self._neg_noise(x.dtype)
        inner(x, u, tmp, ncols=ncols, done_flag=done_flag, ep_coef=ns, ep_add=x)     # u = x - s Q x
        inner(u, w, tmp, ncols=ncols, done_flag=done_flag, ep_coef=ns, ep_add=x)     # w = x - s Q u
        return inner(w, out, tmp, dot_with=dot_with, dot_out=dot_out, ncols=ncols, done_flag=done_flag, ep_coef=ep_coef, ep_add=ep_add)
