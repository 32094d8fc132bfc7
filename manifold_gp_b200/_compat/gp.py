"""Minimal stand-ins for the parts of GPyTorch the reference's kernel / model layer builds on
(``gpytorch.kernels.Kernel / ScaleKernel``, ``gpytorch.models.ExactGP``, ``gpytorch.means.ConstantMean``,
``gpytorch.likelihoods.GaussianLikelihood``, ``gpytorch.distributions.MultivariateNormal``, ``gpytorch.constraints``).

GPyTorch is not installed in this image (nor on the GPU box, and there is no network), so the reference-facing classes in
``manifold_gp_b200.kernels`` / ``manifold_gp_b200.models`` derive from these when ``import gpytorch`` fails and from the
real classes when it succeeds.  Parameter names, shapes and the ``state_dict`` layout follow GPyTorch, so the reference's
shipped checkpoints (``models/*.pth``: ``likelihood.noise_covar.raw_noise``, ``covar_module.raw_outputscale``,
``covar_module.base_kernel.raw_lengthscale``, ``covar_module.base_kernel.raw_graphbandwidth`` + constraint bounds) load
unchanged.  Only what the reference's call sites use is implemented (riemann_kernel.py, riemann_gp.py, train_model.py,
test_model.py).
"""
from __future__ import annotations

import math

import torch
from torch import nn

from .linear_operator import LocalLinearOperator

try:  # pragma: no cover - not installed in this image
    import gpytorch as _gpytorch  # noqa: F401
    HAVE_GPYTORCH = True
except Exception:
    HAVE_GPYTORCH = False


# ---- constraints -----------------------------------------------------------------------------------------------------------
def _inv_softplus(x):
    return x + torch.log(-torch.expm1(-x))


class Interval(nn.Module):
    def __init__(self, lower_bound, upper_bound):
        super().__init__()
        self.register_buffer("lower_bound", torch.as_tensor(float(lower_bound)))
        self.register_buffer("upper_bound", torch.as_tensor(float(upper_bound)))
        # decided once on the host: reading the (device) buffer on every parameter access would synchronise the training loop
        self._unbounded_above = math.isinf(float(upper_bound))

    def transform(self, raw):
        lo, hi = self.lower_bound.to(raw), self.upper_bound.to(raw)
        if self._unbounded_above:
            return torch.nn.functional.softplus(raw) + lo
        return lo + (hi - lo) * torch.sigmoid(raw)

    def inverse_transform(self, value):
        lo, hi = self.lower_bound.to(value), self.upper_bound.to(value)
        if self._unbounded_above:
            return _inv_softplus(value - lo)
        t = (value - lo) / (hi - lo)
        return torch.log(t) - torch.log1p(-t)


class GreaterThan(Interval):
    def __init__(self, lower_bound):
        super().__init__(lower_bound, math.inf)


class Positive(GreaterThan):
    def __init__(self):
        super().__init__(0.0)


# ---- module base ----------------------------------------------------------------------------------------------------------
class Module(nn.Module):
    def __init__(self):
        super().__init__()
        self._priors = {}

    def register_constraint(self, param_name, constraint):
        self.add_module(param_name + "_constraint", constraint)

    def register_prior(self, name, prior, param_or_closure, setting_closure=None):
        self.add_module(name, prior) if isinstance(prior, nn.Module) else None
        self._priors[name] = (prior, param_or_closure, setting_closure)

    def named_priors(self, memo=None, prefix=""):
        for mname, module in self.named_modules():
            pri = getattr(module, "_priors", None)
            if pri:
                for name, (prior, closure, setting) in pri.items():
                    yield (mname + "." if mname else "") + name, module, prior, closure, setting

    def initialize(self, **kwargs):
        for name, val in kwargs.items():
            if "." in name:
                head, tail = name.split(".", 1)
                getattr(self, head).initialize(**{tail: val})
            elif name in self._parameters:
                p = self._parameters[name]
                with torch.no_grad():
                    p.copy_(torch.as_tensor(val, dtype=p.dtype, device=p.device).expand_as(p))
            else:
                with torch.no_grad():
                    setattr(self, name, val)
        return self


# ---- operators the kernel returns (linear_operator's root family) ---------------------------------------------------------
class MatmulLinearOperator(LocalLinearOperator):
    def __init__(self, left, right):
        super().__init__(left, right)
        self.left, self.right = left, right

    def _matmul(self, rhs):
        return self.left @ (self.right @ rhs)

    def _size(self):
        return torch.Size([self.left.shape[-2], self.right.shape[-1]])

    def _transpose_nonbatch(self):
        return MatmulLinearOperator(self.right.transpose(-1, -2), self.left.transpose(-1, -2))

    def to_dense(self):
        return self.left @ self.right

    def _diagonal(self):
        return (self.left * self.right.transpose(-1, -2)).sum(-1)

    def evaluate_kernel(self):
        return self

    def mul(self, s):
        return MatmulLinearOperator(self.left * s, self.right)

    def add_diagonal(self, d):
        return DensePlusDiag(self, d)


class RootLinearOperator(MatmulLinearOperator):
    def __init__(self, root):
        super().__init__(root, root.transpose(-1, -2))
        self.root = root

    def mul(self, s):
        return type(self)(self.root * torch.as_tensor(s).sqrt())


class LowRankRootLinearOperator(RootLinearOperator):
    pass


class DensePlusDiag(LocalLinearOperator):
    """K + diag(d) (what GaussianLikelihood adds); dense-solve helpers for the small posterior covariances."""

    def __init__(self, base, d):
        super().__init__(base.to_dense() if hasattr(base, "to_dense") else base, d)
        self.base, self.d = base, d

    def to_dense(self):
        b = self.base.to_dense() if hasattr(self.base, "to_dense") else self.base
        return b + torch.diag_embed(self.d.expand(b.shape[-1]))

    def _matmul(self, rhs):
        return self.to_dense() @ rhs

    def _size(self):
        return self.base.shape

    def _transpose_nonbatch(self):
        return self

    def evaluate_kernel(self):
        return self


def _as_dense(op):
    return op.to_dense() if hasattr(op, "to_dense") else op


class DenseOperator(LocalLinearOperator):
    def __init__(self, t):
        super().__init__(t)
        self.t = t

    def _matmul(self, rhs):
        return self.t @ rhs

    def _size(self):
        return self.t.shape

    def _transpose_nonbatch(self):
        return DenseOperator(self.t.transpose(-1, -2))

    def to_dense(self):
        return self.t

    def evaluate_kernel(self):
        return self

    def __iadd__(self, other):
        self.t = self.t + _as_dense(other)
        return self

    def __add__(self, other):
        return DenseOperator(self.t + _as_dense(other))

    def inv_quad_logdet(self, inv_quad_rhs=None, logdet=False, reduce_inv_quad=True):
        chol = torch.linalg.cholesky(self.t)
        iq = torch.zeros((), dtype=self.t.dtype, device=self.t.device)
        if inv_quad_rhs is not None:
            sol = torch.linalg.solve_triangular(chol, inv_quad_rhs, upper=False)
            iq = sol.pow(2).sum(-2)
            if reduce_inv_quad:
                iq = iq.sum(-1)
        ld = chol.diagonal().log().sum() * 2 if logdet else torch.zeros((), dtype=self.t.dtype, device=self.t.device)
        return iq, ld


# ---- distributions / means / likelihood ---------------------------------------------------------------------------------
class MultivariateNormal:
    def __init__(self, mean, covariance):
        self.loc = mean
        self._covar = covariance

    @property
    def mean(self):
        return self.loc

    @property
    def lazy_covariance_matrix(self):
        c = self._covar
        return c if isinstance(c, LocalLinearOperator) else DenseOperator(c)

    @property
    def covariance_matrix(self):
        return _as_dense(self._covar)

    @property
    def variance(self):
        c = self._covar
        return c.diagonal() if isinstance(c, LocalLinearOperator) else torch.diagonal(c, dim1=-2, dim2=-1)

    @property
    def stddev(self):
        return self.variance.clamp_min(1e-12).sqrt()


class ConstantMean(Module):
    def __init__(self):
        super().__init__()
        self.register_parameter("raw_constant", nn.Parameter(torch.zeros(())))

    @property
    def constant(self):
        return self.raw_constant

    def forward(self, x):
        return self.raw_constant.expand(x.shape[:-1])


class _HomoskedasticNoise(Module):
    def __init__(self, noise_constraint=None):
        super().__init__()
        self.register_parameter("raw_noise", nn.Parameter(torch.zeros(1)))
        self.register_constraint("raw_noise", noise_constraint if noise_constraint is not None else GreaterThan(1e-4))

    @property
    def noise(self):
        return self.raw_noise_constraint.transform(self.raw_noise)

    @noise.setter
    def noise(self, value):
        value = torch.as_tensor(value).to(self.raw_noise)
        self.initialize(raw_noise=self.raw_noise_constraint.inverse_transform(value))


class GaussianLikelihood(Module):
    def __init__(self, noise_constraint=None):
        super().__init__()
        self.noise_covar = _HomoskedasticNoise(noise_constraint)

    @property
    def noise(self):
        return self.noise_covar.noise

    @noise.setter
    def noise(self, value):
        self.noise_covar.noise = value

    def forward(self, dist):
        c = dist.lazy_covariance_matrix
        return MultivariateNormal(dist.mean, DenseOperator(c.to_dense() + torch.diag_embed(self.noise.expand(c.shape[-1]))))

    __call__ = forward


# ---- kernels ---------------------------------------------------------------------------------------------------------------
class Kernel(Module):
    has_lengthscale = False

    def __init__(self, batch_shape=torch.Size([]), lengthscale_prior=None, lengthscale_constraint=None, **kwargs):
        super().__init__()
        self._batch_shape = batch_shape
        if self.has_lengthscale:
            self.register_parameter("raw_lengthscale", nn.Parameter(torch.zeros(*batch_shape, 1, 1)))
            self.register_constraint("raw_lengthscale", lengthscale_constraint if lengthscale_constraint is not None else Positive())

    @property
    def batch_shape(self):
        return self._batch_shape

    @property
    def lengthscale(self):
        return self.raw_lengthscale_constraint.transform(self.raw_lengthscale) if self.has_lengthscale else None

    @lengthscale.setter
    def lengthscale(self, value):
        value = torch.as_tensor(value).to(self.raw_lengthscale)
        self.initialize(raw_lengthscale=self.raw_lengthscale_constraint.inverse_transform(value))

    def __call__(self, x1, x2=None, diag=False, **params):
        if x1.dim() == 1:
            x1 = x1.unsqueeze(-1)
        if x2 is None:
            x2 = x1
        elif x2.dim() == 1:
            x2 = x2.unsqueeze(-1)
        return self.forward(x1, x2, diag=diag, **params)


class ScaleKernel(Kernel):
    def __init__(self, base_kernel, outputscale_prior=None, outputscale_constraint=None, **kwargs):
        super().__init__(**kwargs)
        self.base_kernel = base_kernel
        self.register_parameter("raw_outputscale", nn.Parameter(torch.zeros(())))
        self.register_constraint("raw_outputscale", outputscale_constraint if outputscale_constraint is not None else Positive())

    @property
    def outputscale(self):
        return self.raw_outputscale_constraint.transform(self.raw_outputscale)

    @outputscale.setter
    def outputscale(self, value):
        value = torch.as_tensor(value).to(self.raw_outputscale)
        self.initialize(raw_outputscale=self.raw_outputscale_constraint.inverse_transform(value))

    def forward(self, x1, x2, diag=False, **params):
        k = self.base_kernel.forward(x1, x2, diag=diag, **params)
        if diag:
            return k * self.outputscale
        return k.mul(self.outputscale) if hasattr(k, "mul") and isinstance(k, LocalLinearOperator) else k * self.outputscale


# ---- exact GP ---------------------------------------------------------------------------------------------------------------
class ExactGP(Module):
    """Train mode: ``model(x)`` is the prior at x.  Eval mode: ``model(x)`` is the exact posterior at x given the training
    data (mean / covariance by the Woodbury identity on the kernel's low-rank root when it has one, dense otherwise)."""

    def __init__(self, train_inputs, train_targets, likelihood):
        super().__init__()
        self.train_inputs = (train_inputs,) if torch.is_tensor(train_inputs) else tuple(train_inputs)
        self.train_targets = train_targets
        self.likelihood = likelihood

    def __call__(self, *inputs):
        x = inputs[0]
        if self.training:
            return self.forward(x)
        tx = self.train_inputs[0]
        prior_train = self.forward(tx)
        mean_train = prior_train.mean
        resid = (self.train_targets - mean_train).unsqueeze(-1)
        noise = self.likelihood.noise
        kern = self.covar_module
        k_tt = prior_train.lazy_covariance_matrix
        k_st = kern(x if x.dim() > 1 else x.unsqueeze(-1), tx)                 # test x train
        k_ss = kern(x if x.dim() > 1 else x.unsqueeze(-1))
        mean_s = self.mean_module(x if x.dim() > 1 else x.unsqueeze(-1))
        root = getattr(k_tt, "root", None)
        if root is not None and root.shape[-1] < root.shape[-2]:
            # (Z Z^T + s I)^-1 b = (b - Z (s I + Z^T Z)^-1 Z^T b) / s
            z = root
            core = z.transpose(-1, -2) @ z + noise * torch.eye(z.shape[-1], dtype=z.dtype, device=z.device)
            chol = torch.linalg.cholesky(core)

            def solve(b):
                return (b - z @ torch.cholesky_solve(z.transpose(-1, -2) @ b, chol)) / noise
        else:
            a = k_tt.to_dense() + noise * torch.eye(k_tt.shape[-1], dtype=resid.dtype, device=resid.device)
            chol = torch.linalg.cholesky(a)

            def solve(b):
                return torch.cholesky_solve(b, chol)
        kst = k_st.to_dense()
        mean = mean_s + (kst @ solve(resid)).squeeze(-1)
        covar = k_ss.to_dense() - kst @ solve(kst.transpose(-1, -2))
        return MultivariateNormal(mean, DenseOperator(covar))
