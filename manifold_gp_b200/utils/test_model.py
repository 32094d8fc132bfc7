"""``test_model`` -- manifold_gp/utils/test_model.py:10-30.

Scores a trained model on held-out points: the posterior (geometric kernel, optionally blended with a Euclidean base model)
is evaluated once, then

    rmse = sqrt(mean((y - mean)^2))
    nll  = 1/2 [ e^T K^-1 e + log|K| + m log 2 pi ] / m        with e = y - mean, K the posterior covariance of the m points

where ``inv_quad_logdet`` of the posterior covariance takes the Cholesky branch for m <= ``max_cholesky`` and the CUDA
mBCG + stochastic-Lanczos branch beyond it.  The settings contexts are the ones the reference enters (fast_pred_var,
max_cholesky_size, cg_tolerance, max_cg_iterations), from ``manifold_gp_b200.settings`` (gpytorch's when installed).
"""
from __future__ import annotations

import math

import torch

from .. import settings


def test_model(model, input, output, noisy_test=False, base_model=None, max_cholesky=800, cg_tolerance=1e-2, cg_iterations=1000):
    with torch.no_grad(), settings.fast_pred_var(), settings.max_cholesky_size(max_cholesky), \
            settings.cg_tolerance(cg_tolerance), settings.max_cg_iterations(cg_iterations):
        model.likelihood.eval()
        model.eval()
        if base_model is not None:
            model.posterior(input, noisy_posterior=noisy_test, base_model=base_model)
        else:
            model.posterior(input, noisy_posterior=noisy_test)
        error = output - model.posterior_mean
        rmse = (error.square().mean()).sqrt()
        inv_quad, logdet = model.posterior_covar.inv_quad_logdet(inv_quad_rhs=error.unsqueeze(-1), logdet=True)
        nll = 0.5 * sum([inv_quad, logdet, error.size(-1) * math.log(2 * math.pi)]) / error.size(-1)
        return rmse, nll


test_model.__test__ = False   # not a pytest test
