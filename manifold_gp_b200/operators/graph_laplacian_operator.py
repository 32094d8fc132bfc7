"""``GraphLaplacianOperator`` -- drop-in for manifold_gp/operators/graph_laplacian_operator.py on B200.

Same constructor, attributes and ``LinearOperator`` protocol as the reference; the arithmetic runs in
libmgp_b200.so:

* cached value properties (:52-106)  -> one fused, deterministic value build (``mgp_lap_values``) over the row-major
  directed structure that is built once per graph and cached on ``idx`` (``graph.structure_for``);
* ``_matmul`` (:108-124, two ``torch_sparse.spmm`` + diagonal + D^{+-1/2} scalings) -> ONE fused SpMM launch
  (``mgp_lap_spmm``) with the diagonal and the degree scalings applied on the fly;
* ``out_of_sample`` (:146-157) -> ``mgp_out_of_sample`` (ELL SpMM, no [Q,k,m] temporary);
* ``diagonalization`` (:132-144) -> CUDA Lanczos driver (``solvers.diagonalization``).
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor
from torch.nn.functional import normalize

from .. import _lib, graph, settings
from .._compat.linear_operator import LinearOperator
from .._lib import c_int32, c_int64, ptr, stream


class GraphLaplacianOperator(LinearOperator):
    def __init__(
            self,
            x: Tensor,                    # squared kNN distances of the M upper-triangular edges
            idx: Tensor,                  # [2, M], row < col
            operator_dimension: int,
            graphbandwidth: Tensor,
            normalization: Optional[str] = "randomwalk",  # "symmetric"
            self_loops: Optional[bool] = True,
            transposed: Optional[bool] = False,
    ):
        super().__init__(
            x,
            idx=idx,
            operator_dimension=operator_dimension,
            graphbandwidth=graphbandwidth,
            normalization=normalization,
            self_loops=self_loops,
            transposed=transposed,
        )
        self.x = x
        self.idx = idx
        self.operator_dimension = int(operator_dimension)
        self.graphbandwidth = graphbandwidth
        self.normalization = normalization
        self.self_loops = self_loops
        self.transposed = transposed
        self._mgp_cache = {}

    # ---- structure / values -----------------------------------------------------------------------------------------
    @property
    def structure(self) -> graph.GraphStructure:
        st = self._mgp_cache.get("structure")
        if st is None:
            st = graph.structure_for(self.idx, self.operator_dimension)
            self._mgp_cache["structure"] = st
        return st

    def _grad_mode(self) -> bool:
        """True when results must carry the autograd graph back to ``graphbandwidth``."""
        eps = self.graphbandwidth
        return torch.is_grad_enabled() and torch.is_tensor(eps) and eps.requires_grad

    def _values(self):
        """(deg_unnorm, deg, diag, a_csr) for the current bandwidth; differentiable w.r.t. ``graphbandwidth``.

        Two cache slots: the differentiable build (grad mode, bandwidth requires grad) and the plain one.  A result built
        under ``no_grad`` (the CG / Lanczos drivers) is never handed to a differentiable caller -- that silently dropped
        the bandwidth gradient of ``solve`` / ``inv_quad_logdet`` / ``logdet`` on a fresh operator."""
        if self._grad_mode():
            v = self._mgp_cache.get("values_grad")
            if v is None:
                from ..autograd import lap_values_autograd
                st = self.structure
                v = lap_values_autograd(st, st.d2csr(self.x), self.graphbandwidth, bool(self.self_loops))
                self._mgp_cache["values_grad"] = v
                self._mgp_cache.setdefault("values", tuple(t.detach() for t in v))   # same memory: one value build serves both
            return v
        v = self._mgp_cache.get("values")
        if v is None:
            st = self.structure
            v = graph.lap_values(st, st.d2csr(self.x), self.graphbandwidth, bool(self.self_loops))
            self._mgp_cache["values"] = v
        return v

    def _memo(self, name, fn):
        # results derived from the value build exist per grad mode, like the values themselves
        key = name + "#g" if self._grad_mode() else name
        if key not in self._mgp_cache:
            self._mgp_cache[key] = fn()
        return self._mgp_cache[key]

    # ---- the reference's cached properties (:52-106) -----------------------------------------------------------------
    @property
    def adjacency_unnorm_mat(self) -> Tensor:
        return self._memo("W", lambda: self.x.div(-4 * self.graphbandwidth.square()).exp().squeeze())

    def _ext(self, name, i):
        """i-th value array in the caller's node order (the structure may be internally permuted)."""
        return self._memo(name, lambda: self.structure.to_external(self._values()[i]))

    @property
    def degree_unnorm_mat(self) -> Tensor:
        return self._ext("Dt_ext", 0)

    @property
    def adjacency_mat(self) -> Tensor:
        dt = self.degree_unnorm_mat
        return self._memo("A", lambda: self.adjacency_unnorm_mat.div(dt[self.idx[0, :]] * dt[self.idx[1, :]]))

    @property
    def degree_mat(self) -> Tensor:
        return self._ext("D_ext", 1)

    @property
    def laplacian_diag(self) -> Tensor:
        return self._ext("diag_ext", 2)

    def _diagonal(self) -> Tensor:
        return self.laplacian_diag

    @property
    def laplacian_triu(self) -> Tensor:
        return self._memo("triu", lambda: self._values()[3][self.structure.upper_pos()])

    # internal (structure-order) D^{+-1/2}
    @property
    def _sqrt_degree(self):
        return self._memo("sqrtD", lambda: self._values()[1].sqrt())

    @property
    def _rsqrt_degree(self):
        return self._memo("rsqrtD", lambda: self._values()[1].pow(-0.5))

    # ---- matvec (:108-124) -------------------------------------------------------------------------------------------
    def _pre_post(self):
        if self.normalization == "randomwalk":
            return (self._rsqrt_degree, self._sqrt_degree) if self.transposed else (self._sqrt_degree, self._rsqrt_degree)
        return None, None

    def _matmul(self, rhs: Tensor) -> Tensor:
        squeeze = rhs.dim() == 1
        vec = rhs.unsqueeze(-1) if squeeze else rhs
        if not vec.is_cuda:
            raise RuntimeError("GraphLaplacianOperator._matmul: rhs must be a CUDA tensor (no CPU fallback exists)")
        _, _, diag, a = self._values()
        if vec.dtype != a.dtype:
            vec = vec.to(a.dtype)
        pre, post = self._pre_post()
        from ..autograd import lap_spmm_apply
        st = self.structure
        out = lap_spmm_apply(st, a, diag, vec.contiguous(), None, pre, post, x_external=True, y_external=True)
        return out.squeeze(-1) if squeeze else out

    # ---- fused path used by the CUDA CG / Lanczos drivers (no autograd, caller-owned buffers) ---------------------
    def _native(self) -> bool:
        return True

    def _mgp_structure(self):
        return self.structure

    def _mgp_matvec(self, x: Tensor, out: Tensor, tmp=None, dot_with=None, dot_out=None, ncols=None, done_flag=None,
                    ep_coef=None, ep_add=None):
        if ncols is not None:
            x, out = x[:, :ncols], out[:, :ncols]
            if ep_add is not None:
                ep_add = ep_add[:, :ncols]
        with torch.no_grad():
            _, _, diag, a = self._values()
            pre, post = self._pre_post()
            graph.lap_spmm(self.structure, a.detach(), diag.detach(), x, pre=pre, post=post, out=out,
                           dot_with=dot_with, dot_out=dot_out, done_flag=done_flag, ep_coef=ep_coef, ep_add=ep_add)
        return out

    def _size(self):
        return torch.Size([self.operator_dimension, self.operator_dimension])

    def _transpose_nonbatch(self):
        if self.normalization == "randomwalk":
            return GraphLaplacianOperator(self.x, self.idx, self.operator_dimension, self.graphbandwidth, self.normalization,
                                          self.self_loops, not self.transposed)
        return self

    def _symmetric_twin(self):
        tw = GraphLaplacianOperator(self.x, self.idx, self.operator_dimension, self.graphbandwidth, "symmetric", self.self_loops)
        for k in ("values", "values_grad"):   # the value build does not depend on the normalisation
            if k in self._mgp_cache:
                tw._mgp_cache[k] = self._mgp_cache[k]
        return tw

    # ---- eigendecomposition (:132-144) -------------------------------------------------------------------------------
    def diagonalization(self, method: Optional[str] = None, num_modes: Optional[int] = None):
        from .. import solvers
        n = self.shape[0]
        size = 3 * num_modes if num_modes is not None and 3 * num_modes <= n else n
        with settings.max_root_decomposition_size(size):
            if self.normalization == "symmetric":
                evals, evecs = solvers.diagonalization(self, method=method)
                evals[0] = 0.0
                if num_modes is not None and num_modes < n:
                    evals, evecs = evals[:num_modes], evecs[:, :num_modes]
                return evals, evecs.to_dense()
            evals, evecs = self._symmetric_twin().diagonalization(method, num_modes)
            evecs = evecs * self.degree_mat.pow(-0.5).view(-1, 1)
            evecs = normalize(evecs, p=2, dim=0)
            return evals, evecs

    # ---- Nystrom extension (:146-157) --------------------------------------------------------------------------------
    def out_of_sample(self, x: Tensor, edge_value: Tensor, edge_idx: Tensor) -> Tensor:
        if not (x.is_cuda and edge_value.is_cuda and edge_idx.is_cuda):
            raise RuntimeError("out_of_sample: tensors must be CUDA tensors (no CPU fallback exists)")
        dt = x.dtype
        deg_un, deg = self.degree_unnorm_mat, self.degree_mat
        phi = x if x.stride(1) == 1 else x.contiguous()
        ev = edge_value.to(dt).contiguous()
        ei = edge_idx.to(torch.int64).contiguous()
        nq, k = ev.shape
        m = phi.shape[1]
        out = torch.empty((nq, m), dtype=dt, device=x.device)
        eps = graph._device_scalar(self.graphbandwidth, dt, x.device)
        _lib.call("mgp_out_of_sample_" + _lib.suffix(dt), ptr(ev), ptr(ei), c_int64(nq), c_int32(k), ptr(eps),
                  ptr(deg_un.to(dt)), ptr(deg.to(dt)), c_int32(0 if self.normalization == "symmetric" else 1),
                  ptr(phi), c_int64(phi.stride(0)), c_int32(m), ptr(out), c_int64(out.stride(0)), stream())
        return out
