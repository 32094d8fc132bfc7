"""``RiemannGP`` -- drop-in for manifold_gp/models/riemann_gp.py: ExactGP whose training objective is the precision-form
marginal likelihood (``precision()`` composes Schur / Scale / Noise wrappers over the CUDA Matern precision operator) and
whose prediction is the low-rank spectral kernel, optionally blended with a Euclidean base model by 1 - bump."""
from __future__ import annotations

import torch

from .._compat import gp as _gp
from ..operators import NoiseWrapperOperator, ScaleWrapperOperator, SchurComplementOperator
from ..utils import bump_function

if _gp.HAVE_GPYTORCH:  # pragma: no cover
    import gpytorch
    _ExactGP, _ConstantMean, _MVN = gpytorch.models.ExactGP, gpytorch.means.ConstantMean, gpytorch.distributions.MultivariateNormal
else:
    _ExactGP, _ConstantMean, _MVN = _gp.ExactGP, _gp.ConstantMean, _gp.MultivariateNormal


class RiemannGP(_ExactGP):
    def __init__(self, train_x, train_y, likelihood, kernel, labeled=None):
        super().__init__(train_x, train_y, likelihood)
        self.mean_module = _ConstantMean()
        self.covar_module = kernel
        self.labeled = labeled

    def eval(self):
        self.base_kernel.eval()
        return super().eval()

    def forward(self, x):
        return _MVN(self.mean_module(x), self.covar_module(x))

    def precision(self, noise=True):                                                     # :32-39
        opt = self.base_kernel.precision()
        if self.labeled is not None:
            opt = SchurComplementOperator(opt, self.labeled)
        if hasattr(self.covar_module, 'outputscale'):
            opt = ScaleWrapperOperator(opt, self.covar_module.outputscale)              # sic: multiplies (Appendix C.5)
        if noise:
            opt = NoiseWrapperOperator(opt, self.likelihood.noise)
        return opt

    def modulation(self, x):                                                             # :41-43
        edge_value, _ = self.base_kernel.knn.search(x, 1)
        return bump_function(edge_value.sqrt().squeeze(), self.base_kernel.bump_scale * self.base_kernel.graphbandwidth.squeeze(),
                             self.base_kernel.bump_decay)

    def posterior(self, x, noisy_posterior=False, base_model=None):                      # :45-50
        self.posterior_geom = self.likelihood(self(x)) if noisy_posterior else self(x)
        if base_model is not None:
            self.posterior_base = base_model.likelihood(base_model(x)) if noisy_posterior else base_model(x)
            self.base_scale = 1 - self.modulation(x)
        return self

    @property
    def base_kernel(self):
        return self.covar_module.base_kernel if hasattr(self.covar_module, 'base_kernel') else self.covar_module

    @property
    def posterior_mean(self):
        mean = self.posterior_geom.mean
        if hasattr(self, "posterior_base"):
            mean = mean + self.base_scale * self.posterior_base.mean
        return mean

    @property
    def posterior_covar(self):
        covar = self.posterior_geom.lazy_covariance_matrix.evaluate_kernel()
        if hasattr(self, "posterior_base"):
            covar = covar + torch.outer(self.base_scale, self.base_scale) * \
                self.posterior_base.lazy_covariance_matrix.evaluate_kernel().to_dense()
        return covar

    @property
    def posterior_stddev(self):
        stddev = self.posterior_geom.stddev
        if hasattr(self, "posterior_base"):
            stddev = stddev + self.base_scale * self.posterior_base.stddev
        return stddev
