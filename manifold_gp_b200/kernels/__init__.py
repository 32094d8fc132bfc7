"""Riemann kernels (``manifold_gp.kernels`` surface): the kNN graph, the graph-bandwidth parameter and the spectral feature map
live in ``RiemannKernel``; ``RiemannMaternKernel`` adds the Matern spectral density and precision operator."""
from .riemann_matern_kernel import RiemannMaternKernel
from .riemann_kernel import RiemannKernel

__all__ = ("RiemannKernel", "RiemannMaternKernel")
