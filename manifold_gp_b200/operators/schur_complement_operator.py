"""``SchurComplementOperator`` -- manifold_gp/operators/schur_complement_operator.py:
Q_xx - Q_xz Q_zz^-1 Q_zx on the labelled rows; the inner solve on the unlabelled block is a CG consumer."""
from __future__ import annotations

import torch
from torch import Tensor

from .._compat.linear_operator import LinearOperator


class MaskedOperator(LinearOperator):
    """``linear_operator.operators.MaskedLinearOperator(base, row_mask, col_mask)``: zero-pad on col_mask, apply base,
    select row_mask rows."""

    def __init__(self, base: LinearOperator, row_mask: Tensor, col_mask: Tensor):
        super().__init__(base, row_mask, col_mask)
        self.base = base
        self.row_mask = row_mask
        self.col_mask = col_mask
        self._rows = int(row_mask.sum())
        self._cols = int(col_mask.sum())

    def _matmul(self, rhs):
        squeeze = rhs.dim() == 1
        r = rhs.unsqueeze(-1) if squeeze else rhs
        full = torch.zeros(self.base.shape[1], r.shape[1], dtype=r.dtype, device=r.device)
        full[self.col_mask] = r
        out = self.base._matmul(full)[self.row_mask]
        return out.squeeze(-1) if squeeze else out

    def _size(self):
        return torch.Size([self._rows, self._cols])

    def _transpose_nonbatch(self):
        return MaskedOperator(self.base._transpose_nonbatch(), self.col_mask, self.row_mask)


class SchurComplementOperator(LinearOperator):
    def __init__(self, base: LinearOperator, mask: Tensor):
        super().__init__(base, mask)
        self.base = base
        self.mask = mask

    def _matmul(self, rhs):
        mask = self.mask.to(rhs.device)   # the reference builds its all-ones mask on the CPU (:27, Appendix C.6)
        ones = torch.ones(self.base.shape[0], dtype=torch.bool, device=rhs.device)
        tmp = MaskedOperator(self.base, ones, mask)._matmul(rhs.contiguous())
        out = MaskedOperator(self.base, ~mask, ~mask).solve(tmp[~mask])
        out = MaskedOperator(self.base, mask, ~mask)._matmul(out)
        return tmp[mask] - out

    def _size(self):
        m = int(self.mask.sum())
        return torch.Size([m, m])

    def _transpose_nonbatch(self):
        return SchurComplementOperator(self.base._transpose_nonbatch(), self.mask)
