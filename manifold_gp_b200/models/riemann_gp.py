"""``RiemannGP`` -- the model surface of manifold_gp/models/riemann_gp.py on top of the CUDA operators.

Two jobs, as in the reference:

* **training objective** -- ``precision()`` assembles the operator whose precision-form marginal likelihood
  ``manifold_informed_train`` minimises: Matern precision (CUDA SpMM chain) -> Schur complement onto the labelled points
  (semi-supervised case) -> output-scale wrapper -> 3-term noise wrapper.  Every product with it, every CG / SLQ solve and
  the backward pass run in the kernels of ``libmgp_b200``.
* **prediction** -- an exact GP whose kernel is the low-rank spectral Riemann kernel (``kernel.eval()`` computes the
  eigenpairs once); optionally blended with a Euclidean base model that takes over, through ``1 - bump``, away from the
  sampled manifold.

Reference lines are cited per method (file: manifold_gp/models/riemann_gp.py).
"""
from __future__ import annotations

import torch

from .._compat import gp as _gp
from ..operators import NoiseWrapperOperator, ScaleWrapperOperator, SchurComplementOperator
from ..utils import bump_function

if _gp.HAVE_GPYTORCH:  # pragma: no cover - gpytorch is not installed in this image
    import gpytorch
    _ExactGP, _ConstantMean, _MVN = gpytorch.models.ExactGP, gpytorch.means.ConstantMean, gpytorch.distributions.MultivariateNormal
else:
    _ExactGP, _ConstantMean, _MVN = _gp.ExactGP, _gp.ConstantMean, _gp.MultivariateNormal


class RiemannGP(_ExactGP):
    def __init__(self, train_x, train_y, likelihood, kernel, labeled=None):
        """``labeled``: boolean mask over the kernel's graph nodes (semi-supervised) or None (all nodes carry targets) -- :12-21."""
        super().__init__(train_x, train_y, likelihood)
        self.mean_module = _ConstantMean()
        self.covar_module = kernel
        self.labeled = labeled

    # ---- structure -----------------------------------------------------------------------------------------------------
    @property
    def base_kernel(self):
        """The Riemann kernel itself, with or without a ``ScaleKernel`` around it (:52-54)."""
        outer = self.covar_module
        return outer.base_kernel if hasattr(outer, 'base_kernel') else outer

    def eval(self):
        """Prediction mode: the kernel computes its eigenpairs first (:23-25)."""
        self.base_kernel.eval()
        return super().eval()

    def forward(self, x):
        """Prior at ``x`` (:27-30)."""
        return _MVN(self.mean_module(x), self.covar_module(x))

    # ---- training objective ---------------------------------------------------------------------------------------------
    def precision(self, noise=True):
        """Precision operator of the (noisy) prior on the training targets (:32-39)."""
        opt = self.base_kernel.precision()
        if self.labeled is not None:
            opt = SchurComplementOperator(opt, self.labeled)
        if hasattr(self.covar_module, 'outputscale'):
            opt = ScaleWrapperOperator(opt, self.covar_module.outputscale)              # sic: multiplies (Appendix C.5)
        if noise:
            opt = NoiseWrapperOperator(opt, self.likelihood.noise)
        return opt

    # ---- prediction -------------------------------------------------------------------------------------------------------
    def modulation(self, x):
        """bump(distance to the nearest sampled point): 1 on the manifold, 0 beyond ``bump_scale * eps`` (:41-43).  The
        nearest-neighbour query runs on the CUDA kNN search."""
        kern = self.base_kernel
        nearest_d2, _ = kern.knn.search(x, 1)
        support = kern.bump_scale * kern.graphbandwidth.squeeze()
        return bump_function(nearest_d2.sqrt().squeeze(), support, kern.bump_decay)

    def posterior(self, x, noisy_posterior=False, base_model=None):
        """Evaluate (and keep) the posterior at ``x``; with ``base_model`` also the Euclidean model's and the blend weight (:45-50)."""
        geom = self(x)
        self.posterior_geom = self.likelihood(geom) if noisy_posterior else geom
        if base_model is not None:
            base = base_model(x)
            self.posterior_base = base_model.likelihood(base) if noisy_posterior else base
            self.base_scale = 1 - self.modulation(x)
        return self

    def _has_base(self) -> bool:
        return hasattr(self, "posterior_base")

    @property
    def posterior_mean(self):
        """:56-61"""
        mean = self.posterior_geom.mean
        if self._has_base():
            mean = mean + self.base_scale * self.posterior_base.mean
        return mean

    @property
    def posterior_covar(self):
        """:63-68"""
        covar = self.posterior_geom.lazy_covariance_matrix.evaluate_kernel()
        if self._has_base():
            base_cov = self.posterior_base.lazy_covariance_matrix.evaluate_kernel().to_dense()
            covar = covar + torch.outer(self.base_scale, self.base_scale) * base_cov
        return covar

    @property
    def posterior_stddev(self):
        """:70-75"""
        stddev = self.posterior_geom.stddev
        if self._has_base():
            stddev = stddev + self.base_scale * self.posterior_base.stddev
        return stddev
