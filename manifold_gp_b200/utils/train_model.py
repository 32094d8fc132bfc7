"""``manifold_informed_train`` -- manifold_gp/utils/train_model.py:49-109: precision-form marginal likelihood
1/2 [ y^T Q y - log|Q| + n log 2pi ] - sum log p(theta), Adam step, output-scale renormalisation by the average variance.
Every heavy call (matmul, inv_quad_logdet -> Cholesky or CUDA mBCG + SLQ, backward through the CUDA kernels) goes through
the operators of this package; the numerical knobs are the same ``settings`` context managers the reference uses."""
from __future__ import annotations

import math

import torch

from .. import settings


def manifold_informed_train(model, optimizer, max_iter=100, tolerance=1e-2, update_norm=None, num_rand_vec=100, max_cholesky=800,
                            cg_tolerance=1e-2, cg_max_iter=1000, scheduler=None, verbose=False):
    model.train()
    model.likelihood.train()

    def ctx():
        return settings.max_cholesky_size(max_cholesky), settings.cg_tolerance(cg_tolerance), settings.max_cg_iterations(cg_max_iter)

    def renorm(mode):
        a, b, c = ctx()
        with torch.no_grad(), a, b, c:
            av = model.covar_module.base_kernel.precision()._average_variance(num_rand_vec=num_rand_vec)
            if mode == "div":
                model.covar_module.outputscale = model.covar_module.outputscale / av       # :55
            elif mode == "set":
                model.covar_module.outputscale = 1 / av                                     # :100
            else:
                model.covar_module.outputscale = model.covar_module.outputscale * av       # :104

    if hasattr(model.covar_module, 'outputscale'):
        renorm("div")

    epoch = 0
    prev_loss = 1e6
    num_data = model.train_targets.shape[0]
    loss = None
    while epoch <= max_iter:
        optimizer.zero_grad()
        precision_operator = model.precision()
        a, b, c = ctx()
        with a, b, c:
            y = model.train_targets
            loss = 0.5 * sum([torch.dot(y, precision_operator.matmul(y.view(-1, 1)).squeeze()),
                              -precision_operator.inv_quad_logdet(logdet=True)[1],
                              num_data * math.log(2 * math.pi)])                              # :67-69
            loss_ndim = loss.ndim
            for _, module, prior, closure, _ in model.named_priors():
                prior_term = prior.log_prob(closure(module))
                loss = loss - prior_term.view(*prior_term.shape[:loss_ndim], -1).sum(dim=-1)
            loss = loss / num_data
        if verbose:
            msg = [f"Iteration: {epoch}, Loss: {loss.item():0.3f}",
                   f"Noise Variance: {model.likelihood.noise.item():0.3f}"]
            if hasattr(model.covar_module, 'outputscale'):
                msg += [f"Signal Variance: {model.covar_module.outputscale.item():0.3f}"]
            msg += [f"Lengthscale: {model.base_kernel.lengthscale.item():0.3f}, Graphbandwidth: {model.base_kernel.graphbandwidth.item():0.3f}"]
            print(',\t'.join(msg))
        loss.backward()
        optimizer.step()
        if scheduler is not None:
            scheduler.step(loss)
        epoch += 1
        if abs(loss.item() - prev_loss) <= tolerance:
            break
        if update_norm is not None and epoch % (update_norm + 1) == 0:
            renorm("set")
    if hasattr(model.covar_module, 'outputscale'):
        renorm("mul")
    return loss.item()
