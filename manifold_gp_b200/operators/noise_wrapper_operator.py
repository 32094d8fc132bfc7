"""``NoiseWrapperOperator`` -- manifold_gp/operators/noise_wrapper_operator.py: Q - s Q^2 + s^2 Q^3 (3 inner matvecs)."""
from __future__ import annotations

from torch import Tensor

from .._compat.linear_operator import LinearOperator


class NoiseWrapperOperator(LinearOperator):
    def __init__(self, operator: LinearOperator, noise: Tensor):
        super().__init__(operator, noise=noise)
        self.operator = operator
        self.noise = noise

    def _matmul(self, rhs):
        rhs = rhs.contiguous()
        op = self.operator._matmul                                           # :21-22
        return op(rhs - self.noise * op(rhs - self.noise * op(rhs)))

    def _size(self):
        return self.operator._size()

    def _transpose_nonbatch(self):
        return NoiseWrapperOperator(self.operator._transpose_nonbatch(), self.noise)   # see ScaleWrapperOperator
