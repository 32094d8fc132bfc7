"""``bump_function`` -- manifold_gp/utils/torch_utils.py:38-41 (trivial elementwise; used by the kernel's out-of-sample
features and RiemannGP.modulation)."""
import torch


def bump_function(x, alpha, beta):
    y = torch.zeros_like(x)
    m = x.abs() < alpha
    y[m] = x[m].square().sub(alpha.square()).pow(-1).mul(beta).exp().div(alpha.square().pow(-1).mul(-beta).exp())
    return y
