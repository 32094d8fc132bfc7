"""GP models (``manifold_gp.models`` surface).  ``VanillaGP`` (a plain gpytorch ExactGP, no data-parallel work) is out of scope."""
from .riemann_gp import RiemannGP

__all__ = ("RiemannGP",)
