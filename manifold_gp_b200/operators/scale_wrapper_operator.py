"""``ScaleWrapperOperator`` -- manifold_gp/operators/scale_wrapper_operator.py: Q*s (default) or Q/s (inverse_scale)."""
from __future__ import annotations

from typing import Optional

from torch import Tensor

from .._compat.linear_operator import LinearOperator


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
