"""Oracle: Riemann-Matern kernel surface (spectral features, eval(), out-of-sample) and the exact-GP posterior.
TEST INFRASTRUCTURE ONLY.

Restates ``manifold_gp/kernels/riemann_kernel.py:79-149``, ``riemann_matern_kernel.py:21-25``,
``manifold_gp/utils/torch_utils.py:38-41`` and the part of gpytorch's ``ExactGP`` eval path the reference
reaches through ``models/riemann_gp.py:45-75`` (third party, restated: mean/covariance of a GP with kernel
``s * Z Z^T`` and Gaussian noise -- standard GP regression algebra, computed here via dense solves).
"""

from __future__ import annotations

import torch

from .operators import LaplacianOracle


def bump_function(x, alpha, beta):
    """``bump_function`` (torch_utils.py:38-41): exp(beta/(x^2-alpha^2)) / exp(-beta/alpha^2) inside |x|<alpha."""
    alpha = torch.as_tensor(alpha, dtype=x.dtype)
    y = torch.zeros_like(x)
    m = x.abs() < alpha
    y[m] = x[m].square().sub(alpha.square()).pow(-1).mul(beta).exp().div(alpha.square().pow(-1).mul(-beta).exp())
    return y


def spectral_density(eigval, nu, lengthscale):
    """``RiemannMaternKernel.spectral_density`` (riemann_matern_kernel.py:21-22)."""
    return (2 * nu / lengthscale ** 2 + eigval).pow(-nu)


def eval_eigenpairs(lap: LaplacianOracle, num_modes: int):
    """``RiemannKernel.eval`` (riemann_kernel.py:117-130): dense eigh of the *symmetric-form* Laplacian
    assembled from ``laplacian_diag`` / ``laplacian_triu``; first ``num_modes``; ``eigval[0]=0``;
    eigvec <- D^-1/2 eigvec, column-normalised (applied for every normalisation, Appendix C.10)."""
    n = lap.n
    dense = torch.zeros(n, n, dtype=lap.val.dtype)
    dense[lap.idx[0], lap.idx[1]] = -lap.laplacian_triu
    dense = dense + dense.T
    dense = dense + torch.diag(lap.laplacian_diag)
    eigval, eigvec = torch.linalg.eigh(dense)
    eigval, eigvec = eigval[:num_modes].clone(), eigvec[:, :num_modes].clone()
    eigval[0] = 0.0
    eigvec = eigvec * lap.degree_mat.pow(-0.5).view(-1, 1)
    eigvec = torch.nn.functional.normalize(eigvec, p=2, dim=0)
    return eigval, eigvec


def features(lap: LaplacianOracle, eigval, eigvec, nu, lengthscale, x_is_train=True,
             edge_value=None, edge_index=None, bump_scale=1.0, bump_decay=0.01):
    """``RiemannKernel.features`` (riemann_kernel.py:132-149).

    Train branch (:133-136): sqrt(S/sum(S) * N) * eigvec.
    New-point branch (:138-149): needs the kNN query result ``edge_value[Q,k]`` (squared distances) and
    ``edge_index[Q,k]`` of the new points against the training set.
    """
    eps = lap.eps
    n = eigvec.shape[0]
    if x_is_train:
        s = spectral_density(eigval, nu, lengthscale)
        s = s / s.sum()
        return (s * n).sqrt() * eigvec
    support = edge_value[:, 0].sqrt() < bump_scale * eps
    feats = torch.zeros(edge_value.shape[0], eigvec.shape[1], dtype=eigvec.dtype)
    if support.sum() != 0:
        s = spectral_density(eigval, nu, lengthscale).div((1 - eps.square() * eigval).square())
        s = s / s.sum()
        s = s * n
        feats[support] = s.sqrt() * lap.out_of_sample(eigvec, edge_value[support], edge_index[support]) * \
            bump_function(edge_value[support, 0].sqrt(), bump_scale * eps, bump_decay).unsqueeze(-1)
    return feats


def low_rank_posterior(z_train, z_test, y, outputscale, noise, mean_const=0.0, noisy=False):
    """Posterior mean / covariance of ``ExactGP`` with ``ScaleKernel(RiemannMaternKernel)`` in eval mode
    (riemann_gp.py:45-50 -> gpytorch ``DefaultPredictionStrategy``): K = s Z Z^T, likelihood noise sigma^2.

    Solved through the m x m Woodbury core:  (s Z Z^T + sig I)^-1 = (I - Z (sig/s I + Z^T Z)^-1 Z^T) / sig.
    """
    m = z_train.shape[1]
    core = torch.eye(m, dtype=z_train.dtype) * (noise / outputscale) + z_train.T @ z_train

    def solve(b):
        return (b - z_train @ torch.linalg.solve(core, z_train.T @ b)) / noise

    resid = (y - mean_const).reshape(-1, 1)
    k_star = outputscale * (z_test @ z_train.T)
    mean = mean_const + (k_star @ solve(resid)).squeeze(-1)
    covar = outputscale * (z_test @ z_test.T) - k_star @ solve(k_star.T)
    if noisy:
        covar = covar + noise * torch.eye(z_test.shape[0], dtype=z_test.dtype)
    return mean, covar
