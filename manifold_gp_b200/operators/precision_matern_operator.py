"""``PrecisionMaternOperator`` -- drop-in for manifold_gp/operators/precision_matern_operator.py on B200.

(2 nu / kappa^2 I + L)^nu as nu chained fused SpMM launches: each reference step ``out <- (out + c L out)/c``
(:28-32, c = kappa^2 / 2nu) is one ``mgp_lap_spmm`` with ``shift = 1/c`` folded into the diagonal; for the
random-walk normalisation the D^{1/2} factors of  D (D^-1/2 (shift + L_sym) D^1/2)^nu = D^1/2 (shift + L_sym)^nu D^1/2
(:34-35) ride on the first / last launch.  The last launch can also emit CG's p^T A p (``_mgp_matvec``).
"""
from __future__ import annotations

import torch
from torch import Tensor

from .. import graph
from .._compat.linear_operator import LinearOperator
from .graph_laplacian_operator import GraphLaplacianOperator


class PrecisionMaternOperator(LinearOperator):
    def __init__(self, laplacian: LinearOperator, nu: int, lengthscale: Tensor):
        super().__init__(laplacian, nu=nu, lengthscale=lengthscale)
        self.laplacian = laplacian
        self.nu = nu
        self.lengthscale = lengthscale

    def _shift(self):
        # 1/c with c = kappa^2/(2 nu)  (:27); stays on the device
        return (2.0 * self.nu) / self.lengthscale.square().reshape(-1)[:1]

    def _shift_const(self, dtype):
        """Detached ``_shift()`` for the solver loops, computed once per lengthscale value (not once per matvec)."""
        ls = self.lengthscale
        key = (ls.data_ptr(), ls._version, dtype)
        hit = getattr(self, "_mgp_shift", None)
        if hit is None or hit[0] != key:
            with torch.no_grad():
                hit = (key, self._shift().detach().to(dtype).contiguous())
            self._mgp_shift = hit
        return hit[1]

    def _mgp_cache_key(self, dtype):
        """Identity of everything ``_mgp_matvec`` reads (memory + version): the CG driver keeps its captured CUDA graph
        while this is unchanged."""
        lap = self.laplacian
        _, _, diag, a = lap._values()
        shift = self._shift_const(a.dtype)
        key = [lap.structure is not None and id(lap.structure), a.data_ptr(), a._version, diag.data_ptr(), diag._version,
               shift.data_ptr(), shift._version, self.nu, lap.normalization]
        if lap.normalization == "randomwalk":
            sq = lap._sqrt_degree
            key += [sq.data_ptr(), sq._version]
        return tuple(key)

    def _native(self) -> bool:
        return isinstance(self.laplacian, GraphLaplacianOperator)

    def _matmul(self, rhs: Tensor) -> Tensor:
        if not self._native():  # arbitrary LinearOperator as Laplacian: the reference's formula verbatim
            out = rhs.contiguous()
            diag = self.lengthscale.square().squeeze() / (2 * self.nu)
            for _ in range(self.nu):
                out = out + diag * self.laplacian._matmul(out)
                out = out / diag
            if getattr(self.laplacian, "normalization", None) == "randomwalk":
                out = out * self.laplacian.degree_mat.view(-1, 1)
            return out
        squeeze = rhs.dim() == 1
        vec = rhs.unsqueeze(-1) if squeeze else rhs
        if not vec.is_cuda:
            raise RuntimeError("PrecisionMaternOperator._matmul: rhs must be a CUDA tensor (no CPU fallback exists)")
        lap = self.laplacian
        _, _, diag, a = lap._values()
        if vec.dtype != a.dtype:
            vec = vec.to(a.dtype)
        from ..autograd import lap_spmm_apply
        shift = self._shift().to(a.dtype)
        rw = lap.normalization == "randomwalk"
        sq = lap._sqrt_degree if rw else None
        st = lap.structure
        out = vec.contiguous()
        for s in range(self.nu):
            pre = sq if (rw and s == 0) else None
            post = sq if (rw and s == self.nu - 1) else None
            # the caller's row order is translated on the fly by the first / last launch (no permutation passes)
            out = lap_spmm_apply(st, a, diag, out, shift, pre, post, x_external=(s == 0), y_external=(s == self.nu - 1))
        return out.squeeze(-1) if squeeze else out

    # ---- fused path used by the CUDA CG / Lanczos drivers (no autograd, caller-owned buffers) -------------------------
    def _mgp_matvec(self, x: Tensor, out: Tensor, tmp: Tensor, dot_with=None, dot_out=None, ncols=None, done_flag=None,
                    ep_coef=None, ep_add=None):
        """out[:, :ncols] <- P x[:, :ncols] on caller-owned [n, ld] buffers (unit column stride); ``tmp`` is scratch of the
        same shape (used when nu > 1).  If ``dot_out`` is given, dot_out[c] = sum_i dot_with[i,c] * out[i,c] comes out of
        the last launch (``dot_with`` must share x's leading dimension).  ``done_flag``: device scalar of the solver state; the
        launches are no-ops once it is non-zero (CG chunks replayed past convergence).  ``ep_coef`` / ``ep_add``: the wrappers'
        algebra on the LAST launch, out <- ep_add + ep_coef * (P x) (``graph.lap_spmm``)."""
        lap = self.laplacian
        st = lap.structure
        if ncols is not None:
            x, out, tmp = x[:, :ncols], out[:, :ncols], (tmp[:, :ncols] if tmp is not None else None)
            if ep_add is not None:
                ep_add = ep_add[:, :ncols]
        with torch.no_grad():
            _, _, diag, a = lap._values()
            shift = self._shift_const(a.dtype)
            rw = lap.normalization == "randomwalk"
            sq = lap._sqrt_degree if rw else None
            src = x
            for s in range(self.nu):
                last = s == self.nu - 1
                dst = out if ((self.nu - 1 - s) % 2 == 0) else tmp
                graph.lap_spmm(st, a.detach(), diag.detach(), src, shift=shift, pre=sq if (rw and s == 0) else None,
                               post=sq if (rw and last) else None, out=dst,
                               dot_with=dot_with if last else None, dot_out=dot_out if last else None, done_flag=done_flag,
                               ep_coef=ep_coef if last else None, ep_add=ep_add if last else None)
                src = dst
        return out

    def _mgp_structure(self):
        return self.laplacian.structure

    def _size(self):
        return self.laplacian._size()

    def _transpose_nonbatch(self):
        return self

    def _average_variance(self, num_rand_vec=100):
        """(1/R) sum_j e_j^T Q^-1 e_j over R random one-hot vectors (:45-53), a batched CG solve."""
        d = self.shape[0]
        dev, dt = self.lengthscale.device, self.dtype
        if num_rand_vec >= d:
            rand_vec = torch.eye(d, device=dev, dtype=dt)
        else:
            rand_idx = torch.randint(0, d - 1, (1, num_rand_vec), device=dev)   # sic: never picks d-1, may repeat (:50)
            rand_vec = torch.zeros(d, num_rand_vec, device=dev, dtype=dt).scatter_(0, rand_idx, 1.0)
        return self.inv_quad_logdet(inv_quad_rhs=rand_vec, logdet=False)[0] / rand_vec.shape[1]
