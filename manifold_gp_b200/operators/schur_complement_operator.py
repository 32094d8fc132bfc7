"""``SchurComplementOperator`` -- manifold_gp/operators/schur_complement_operator.py:
Q_xx - Q_xz Q_zz^-1 Q_zx on the labelled rows; the inner solve on the unlabelled block is a CG consumer."""
from __future__ import annotations

import torch
from torch import Tensor

from .._compat.linear_operator import LinearOperator
from .scale_wrapper_operator import fused_wrappers_enabled


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


class PrincipalBlockOperator(LinearOperator):
    """K Q K + (I - K) on the FULL index space, K = diag(keep): the principal block ``Q[keep, keep]`` padded with an identity.

    The reference solves with ``MaskedLinearOperator(base, ~mask, ~mask)`` on the compressed index space
    (schur_complement_operator.py:28); the same Krylov process runs here on full-size vectors that are zero on the dropped rows
    (right-hand side zero there => residuals, directions and iterates stay zero there, alpha / beta are identical), which lets
    the inner solve use the fused CUDA CG (tile streams, dot epilogue, CUDA graph) of the wrapped operator unchanged: a matvec
    is the base's fused product followed by one masking pass."""

    def __init__(self, base: LinearOperator, keep: Tensor):
        super().__init__(base, keep)
        self.base = base
        self.keep = keep
        self._mgp_keep = {}

    def _matmul(self, rhs):
        squeeze = rhs.dim() == 1
        r = rhs.unsqueeze(-1) if squeeze else rhs
        k = self.keep.to(r.device).unsqueeze(-1)
        kf = k.to(r.dtype)
        out = self.base._matmul((r * kf).contiguous()) * kf + r * (1 - kf)
        return out.squeeze(-1) if squeeze else out

    def _size(self):
        return self.base._size()

    def _transpose_nonbatch(self):
        return PrincipalBlockOperator(self.base._transpose_nonbatch(), self.keep)

    # ---- solver-driver interface (see PrecisionMaternOperator._mgp_matvec) ---------------------------------------------------
    def _native(self) -> bool:
        inner = self.base
        return hasattr(inner, "_mgp_matvec") and getattr(inner, "_native", lambda: True)()

    def _mgp_structure(self):
        return self.base._mgp_structure()

    def _keep_internal(self, dtype):
        """0/1 vector [n, 1] in the structure's (permuted) row order."""
        hit = self._mgp_keep.get(dtype)
        if hit is None:
            st = self._mgp_structure()
            hit = st.to_internal(self.keep.to(st.device).to(dtype).unsqueeze(-1)).contiguous()
            self._mgp_keep[dtype] = hit
        return hit

    def _mgp_cache_key(self, dtype):
        k = self._keep_internal(dtype)
        return ("principal", k.data_ptr(), k._version) + tuple(self.base._mgp_cache_key(dtype))

    def _mgp_matvec(self, x: Tensor, out: Tensor, tmp: Tensor, dot_with=None, dot_out=None, ncols=None, done_flag=None,
                    ep_coef=None, ep_add=None):
        """out <- K Q x for x that is zero on the dropped rows (the CG invariant); the fused dot product of the base launch is
        already the masked one when ``dot_with`` is zero there (it is: CG passes p)."""
        self.base._mgp_matvec(x, out, tmp, dot_with=dot_with, dot_out=dot_out, ncols=ncols, done_flag=done_flag)
        view = out if ncols is None else out[:, :ncols]
        view.mul_(self._keep_internal(out.dtype))
        if ep_coef is not None:              # (not used by the Schur path itself; kept so the interface composes)
            view.mul_(ep_coef.to(out.dtype))
            if dot_out is not None:
                dot_out.mul_(ep_coef.to(out.dtype))
        if ep_add is not None:
            view.add_(ep_add if ncols is None else ep_add[:, :ncols])
        return out


class SchurComplementOperator(LinearOperator):
    def __init__(self, base: LinearOperator, mask: Tensor):
        super().__init__(base, mask)
        self.base = base
        self.mask = mask

    def _matmul(self, rhs):
        mask = self.mask.to(rhs.device)   # the reference builds its all-ones mask on the CPU (:27, Appendix C.6)
        ones = torch.ones(self.base.shape[0], dtype=torch.bool, device=rhs.device)
        tmp = MaskedOperator(self.base, ones, mask)._matmul(rhs.contiguous())
        base = self.base
        if hasattr(base, "_mgp_matvec") and getattr(base, "_native", lambda: True)() and fused_wrappers_enabled():
            # inner solve on the full index space with the fused CUDA CG (PrincipalBlockOperator); same iterates as :28
            keep = ~mask
            z = tmp * keep.unsqueeze(-1).to(tmp.dtype) if tmp.dim() == 2 else tmp * keep.to(tmp.dtype)
            inner = self.__dict__.get("_mgp_inner")          # kept: its CG buffers / captured graph serve every outer matvec
            if inner is None or inner.keep.device != keep.device:
                inner = PrincipalBlockOperator(base, keep)
                self.__dict__["_mgp_inner"] = inner
            sol = inner.solve(z)
            return tmp[mask] - base._matmul(sol.contiguous())[mask]
        out = MaskedOperator(self.base, ~mask, ~mask).solve(tmp[~mask])
        out = MaskedOperator(self.base, mask, ~mask)._matmul(out)
        return tmp[mask] - out

    def _size(self):
        m = int(self.mask.sum())
        return torch.Size([m, m])

    def _transpose_nonbatch(self):
        return SchurComplementOperator(self.base._transpose_nonbatch(), self.mask)
