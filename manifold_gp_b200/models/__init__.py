from .riemann_gp import RiemannGP

__all__ = ["RiemannGP"]
