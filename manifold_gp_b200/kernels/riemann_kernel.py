"""``RiemannKernel`` -- drop-in for manifold_gp/kernels/riemann_kernel.py on B200.

Owns the kNN graph (built once by the CUDA kNN + symmetrise kernels), the graph-bandwidth parameter
``raw_graphbandwidth[1,1]`` (+ ``Positive()`` constraint by default, :48-63) and the spectral feature map.  ``laplacian()``
returns a fresh CUDA ``GraphLaplacianOperator`` per call exactly like the reference (:114-115): the row-major structure is
cached on ``edge_index``, only the O(nnz) value build is redone when the bandwidth changes.

``eval()`` (:117-130): the reference assembles the dense N x N Laplacian and calls ``torch.linalg.eigh`` -- O(N^3), which
makes BASELINE cfg-B (70k points) impossible.  Here the dense branch is kept for ``N <= dense_eigh_limit`` (bit-for-bit the
reference's post-processing) and larger graphs use the Lanczos path the reference left commented out at :120
(``laplacian_operator.diagonalization(num_modes=...)``), with the eigensolver swapped for one that converges at the
bottom of the spectrum (``solvers.smallest_eigenpairs``, Chebyshev-filtered subspace iteration over the CUDA SpMM).
"""
from __future__ import annotations

from abc import abstractmethod
from typing import Optional

import torch
from torch import Tensor
from torch.nn.functional import normalize

from .._compat import gp as _gp
from ..operators import GraphLaplacianOperator
from ..utils import NearestNeighbors, bump_function

if _gp.HAVE_GPYTORCH:  # pragma: no cover - not installed in this image
    import gpytorch
    from gpytorch.constraints import Positive
    from linear_operator.operators import LowRankRootLinearOperator, MatmulLinearOperator, RootLinearOperator
    _KernelBase = gpytorch.kernels.Kernel
    _Prior = gpytorch.priors.Prior
else:
    from .._compat.gp import LowRankRootLinearOperator, MatmulLinearOperator, Positive, RootLinearOperator
    _KernelBase = _gp.Kernel
    _Prior = torch.nn.Module


class RiemannKernel(_KernelBase):
    has_lengthscale = True
    dense_eigh_limit = 20000
    large_graph_method = "chebyshev"      # "lanczos" reproduces linear_operator's diagonalization (see eval())

    def __init__(self,
                 x: torch.Tensor,
                 nearest_neighbors: Optional[int] = 10,
                 laplacian_normalization: Optional[str] = "symmetric",
                 num_modes: Optional[int] = 100,
                 bump_scale: Optional[float] = 1.0,
                 bump_decay: Optional[float] = 0.01,
                 graphbandwidth_prior=None,
                 graphbandwidth_constraint=None,
                 graph_file: Optional[str] = None,
                 **kwargs):
        super(RiemannKernel, self).__init__(**kwargs)

        self.knn = NearestNeighbors(x, nlist=1)
        self.nearest_neighbors = nearest_neighbors
        # graph_file (extension): reuse a graph saved by an earlier run instead of searching again; the
        # file is written after the first build.  None = the reference's behaviour (always rebuild, :40-42).
        import os
        from .. import graph as _graph
        if graph_file is not None and os.path.exists(graph_file):
            self.edge_index, self.edge_value = _graph.load_knn_graph(graph_file, x.device, x=x, k=nearest_neighbors)
            self.edge_value = self.edge_value.to(x.dtype)
        else:
            self.edge_index, self.edge_value = self.knn.graph(self.nearest_neighbors, nprobe=1)
            if graph_file is not None:
                _graph.save_knn_graph(graph_file, self.edge_index, self.edge_value, x.shape[0], nearest_neighbors, x=x)
        self.laplacian_normalization = laplacian_normalization
        self.num_modes = num_modes
        self.bump_scale = bump_scale
        self.bump_decay = bump_decay

        if graphbandwidth_constraint is None:
            graphbandwidth_constraint = Positive()

        self.register_parameter('raw_graphbandwidth', torch.nn.Parameter(torch.zeros(*self.batch_shape, 1, 1)))

        if graphbandwidth_prior is not None:
            if not isinstance(graphbandwidth_prior, _Prior):
                raise TypeError("Expected gpytorch.priors.Prior but got " + type(graphbandwidth_prior).__name__)
            self.register_prior(
                "graphbandwidth_prior", graphbandwidth_prior, self._graphbandwidth_param, self._graphbandwidth_closure
            )

        self.register_constraint("raw_graphbandwidth", graphbandwidth_constraint)

    def _graphbandwidth_param(self, m) -> Tensor:
        return m.graphbandwidth

    def _graphbandwidth_closure(self, m, v: Tensor) -> Tensor:
        return m._set_graphbandwidth(v)

    def _set_graphbandwidth(self, value: Tensor):
        if not torch.is_tensor(value):
            value = torch.as_tensor(value).to(self.raw_graphbandwidth)
        self.initialize(raw_graphbandwidth=self.raw_graphbandwidth_constraint.inverse_transform(value))

    @property
    def graphbandwidth(self) -> Tensor:
        return self.raw_graphbandwidth_constraint.transform(self.raw_graphbandwidth)

    @graphbandwidth.setter
    def graphbandwidth(self, value: Tensor):
        self._set_graphbandwidth(value)

    # ---- kernel matrix as a (low-rank) root operator (:79-100) ----------------------------------------------------------
    def forward(self, x1: Tensor, x2: Tensor, diag: bool = False, last_dim_is_batch: bool = False, **kwargs):
        if last_dim_is_batch:
            x1 = x1.transpose(-1, -2).unsqueeze(-1)
            x2 = x2.transpose(-1, -2).unsqueeze(-1)
        x1_eq_x2 = torch.equal(x1, x2)
        z1 = self.features(x1)
        z2 = z1 if x1_eq_x2 else self.features(x2)
        if diag:
            return (z1 * z2).sum(-1)
        if x1_eq_x2:
            if z1.size(-1) < z2.size(-2):
                return LowRankRootLinearOperator(z1)
            return RootLinearOperator(z1)
        return MatmulLinearOperator(z1, z2.transpose(-1, -2))

    @abstractmethod
    def spectral_density(self):
        raise NotImplementedError()

    def laplacian(self):
        return GraphLaplacianOperator(self.edge_value, self.edge_index, self.knn.x.shape[0], self.graphbandwidth,
                                      self.laplacian_normalization)

    # ---- eigen-decomposition for prediction (:117-130) --------------------------------------------------------------------
    def eval(self):
        self.laplacian_operator = self.laplacian()
        with torch.no_grad():
            n = self.laplacian_operator.operator_dimension
            if n <= self.dense_eigh_limit:
                dense = self.laplacian_operator._symmetric_twin().to_dense()     # diag - S - S^T, as assembled at :121-123
                dense = 0.5 * (dense + dense.T)
                eigval, eigvec = torch.linalg.eigh(dense)
                eigval, eigvec = eigval[:self.num_modes].clone(), eigvec[:, :self.num_modes].clone()
            else:
                # the SMALLEST num_modes eigenpairs: Chebyshev-filtered subspace iteration over the CUDA SpMM
                # (solvers.smallest_eigenpairs).  method="lanczos" -- linear_operator's 3 * num_modes-step Lanczos, the
                # call the reference left commented out at :120 -- converges at the top of the spectrum first and does not
                # deliver them at this size.
                eigval, eigvec = self.laplacian_operator._symmetric_twin().diagonalization(method=self.large_graph_method,
                                                                                           num_modes=self.num_modes)
            eigval[0] = 0.0
            eigvec = eigvec * self.laplacian_operator.degree_mat.pow(-0.5).view(-1, 1)
            self.eigval, self.eigvec = eigval, normalize(eigvec, p=2, dim=0)
        return super().eval()

    # ---- spectral features (:132-149) ---------------------------------------------------------------------------------------
    def features(self, x: Tensor) -> Tensor:
        if x.shape == self.knn.x.shape and torch.equal(x, self.knn.x):
            spectral_density = self.spectral_density()
            spectral_density = spectral_density / spectral_density.sum()
            return (spectral_density * self.eigvec.shape[0]).sqrt() * self.eigvec
        edge_value, edge_index = self.knn.search(x, self.nearest_neighbors)
        x_within_support = edge_value[:, 0].sqrt() < self.bump_scale * self.graphbandwidth.squeeze()
        features = torch.zeros(x.shape[0], self.num_modes, device=x.device, dtype=self.eigvec.dtype)
        if x_within_support.sum() != 0:
            spectral_density = self.spectral_density().div((1 - self.graphbandwidth.square() * self.eigval).square())
            spectral_density = spectral_density / spectral_density.sum()
            spectral_density = spectral_density * self.knn.x.shape[0]
            ext = self.laplacian_operator.out_of_sample(self.eigvec, edge_value[x_within_support], edge_index[x_within_support])
            features[x_within_support] = spectral_density.sqrt() * ext * \
                bump_function(edge_value[x_within_support, 0].sqrt(), self.bump_scale * self.graphbandwidth.squeeze(),
                              self.bump_decay).unsqueeze(-1)
        return features
