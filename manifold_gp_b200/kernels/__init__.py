from .riemann_kernel import RiemannKernel
from .riemann_matern_kernel import RiemannMaternKernel

__all__ = ["RiemannKernel", "RiemannMaternKernel"]
