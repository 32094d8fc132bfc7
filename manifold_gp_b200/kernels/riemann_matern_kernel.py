"""``RiemannMaternKernel`` -- manifold_gp/kernels/riemann_matern_kernel.py.

The Matern member of the Riemann kernel family on the kNN graph held by ``RiemannKernel``:

* ``spectral_density()``  S(lambda) = (2 nu / kappa^2 + lambda)^(-nu) on the Laplacian eigenvalues the kernel computed in
  ``eval()`` -- the weights of the spectral feature map (riemann_matern_kernel.py:21-22);
* ``precision()``         the operator (2 nu / kappa^2 I + L)^nu applied as nu chained fused SpMM launches
  (``PrecisionMaternOperator``; :24-25) -- what the training loss and the CG / SLQ solves consume.

kappa is the kernel's ``lengthscale`` parameter (``has_lengthscale``), eps its ``graphbandwidth``.
"""
from __future__ import annotations

from typing import Optional

from ..operators import PrecisionMaternOperator
from .riemann_kernel import RiemannKernel


class RiemannMaternKernel(RiemannKernel):
    has_lengthscale = True

    def __init__(self, nu: Optional[int] = 2, **kwargs):
        """``nu``: integer smoothness (number of chained Laplacian products); the remaining keyword arguments are
        ``RiemannKernel``'s (x, nearest_neighbors, laplacian_normalization, num_modes, bump_scale, bump_decay, ...)."""
        super().__init__(**kwargs)
        self.nu = nu

    def _shift(self):
        """2 nu / kappa^2 as a tensor that keeps its autograd link to the lengthscale."""
        return 2 * self.nu / self.lengthscale.square()

    def spectral_density(self):
        return (self._shift() + self.eigval).pow(-self.nu)

    def precision(self):
        return PrecisionMaternOperator(self.laplacian(), self.nu, self.lengthscale)
