"""``RiemannMaternKernel`` -- manifold_gp/kernels/riemann_matern_kernel.py: spectral density (2nu/kappa^2 + lambda)^-nu and the
Matern precision operator (2nu/kappa^2 I + L)^nu."""
from __future__ import annotations

from typing import Optional

from ..operators import PrecisionMaternOperator
from .riemann_kernel import RiemannKernel


class RiemannMaternKernel(RiemannKernel):
    has_lengthscale = True

    def __init__(self, nu: Optional[int] = 2, **kwargs):
        super().__init__(**kwargs)
        self.nu = nu

    def spectral_density(self):
        return (2 * self.nu / self.lengthscale.square() + self.eigval).pow(-self.nu)      # :21-22

    def precision(self):
        return PrecisionMaternOperator(self.laplacian(), self.nu, self.lengthscale)       # :24-25
