"""GPU parity of the kernel / model surface against the oracle on the
reference's own fixture and configuration (cfg-A: examples/1D_supervised_learning -- dumbbell curve, nearest_neighbors=10,
random-walk normalisation, nu=1, 50 modes, bump_scale=10, bump_decay=1): RiemannMaternKernel.eval / features / forward,
RiemannGP posterior mean and variance (tolerance 1e-4, BASELINE.json north_star), manifold_informed_train's loss."""
import math

import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu
DEV = "cuda"
K, NU, MODES, BUMP_SCALE, BUMP_DECAY = 10, 1, 50, 10.0, 1.0
EPS, KAPPA, OUTPUTSCALE, NOISE = 0.05, 0.7, 1.7, 1e-2


def _build(dumbbell, dtype, normalization="randomwalk"):
    import manifold_gp_b200 as mgp
    from manifold_gp_b200._compat import gp as gpc
    x = dumbbell["train_x"].to(dtype).to(DEV)
    y = dumbbell["train_y"].to(dtype).to(DEV)
    kernel = mgp.RiemannMaternKernel(nu=NU, x=x, nearest_neighbors=K, laplacian_normalization=normalization,
                                     num_modes=MODES, bump_scale=BUMP_SCALE, bump_decay=BUMP_DECAY).to(DEV).to(dtype)
    kernel.graphbandwidth = torch.tensor([[EPS]], dtype=dtype, device=DEV)
    kernel.lengthscale = torch.tensor([[KAPPA]], dtype=dtype, device=DEV)
    covar = gpc.ScaleKernel(kernel).to(DEV).to(dtype)
    covar.outputscale = torch.tensor(OUTPUTSCALE, dtype=dtype, device=DEV)
    lik = gpc.GaussianLikelihood().to(DEV).to(dtype)
    lik.noise = torch.tensor([NOISE], dtype=dtype, device=DEV)
    model = mgp.RiemannGP(x, y, lik, covar).to(DEV)
    return model, kernel, x, y


def _oracle_side(dumbbell, dtype, normalization="randomwalk"):
    x = dumbbell["train_x"].to(dtype)
    idx, val = oracle.knn_graph(x.float(), K)
    lap = oracle.LaplacianOracle(val.to(dtype), idx, x.shape[0], EPS, normalization, True)
    eigval, eigvec = oracle.eval_eigenpairs(lap, MODES)
    return x, lap, eigval, eigvec


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-8), (torch.float32, 1e-4)])
def test_eval_eigenpairs_and_train_features(dumbbell, dtype, tol):
    model, kernel, x, y = _build(dumbbell, dtype)
    xo, lap, eigval, eigvec = _oracle_side(dumbbell, dtype)
    model.eval()
    ev, evec = kernel.eigval.cpu(), kernel.eigvec.cpu()
    assert ev.shape == (MODES,) and evec.shape == (x.shape[0], MODES)
    assert float((ev - eigval).abs().max() / eigval.abs().max()) < tol
    # the kernel matrix Z Z^T is invariant to the eigenvectors' signs / rotations inside eigenspaces: compare that
    z = kernel.features(x).cpu()
    zo = oracle.features(lap, eigval, eigvec, NU, KAPPA, x_is_train=True)
    kk, ko = z[:300] @ z.T, zo[:300] @ zo.T
    assert float((kk - ko).abs().max() / ko.abs().max()) < (1e-6 if dtype == torch.float64 else 2e-3)


@pytest.mark.parametrize("normalization", ["randomwalk", "symmetric"])
def test_out_of_sample_features_and_posterior(dumbbell, normalization):
    dtype = torch.float64
    model, kernel, x, y = _build(dumbbell, dtype, normalization)
    xo, lap, eigval, eigvec = _oracle_side(dumbbell, dtype, normalization)
    xt = dumbbell["test_x"].to(dtype)
    model.eval()
    model.likelihood.eval()
    # oracle features of the held-out points: kNN query + Nystrom extension + bump (riemann_kernel.py:138-149)
    ev, ei = oracle.knn_search(xo.float(), xt.float(), K)
    zo_tr = oracle.features(lap, eigval, eigvec, NU, KAPPA, x_is_train=True)
    zo_te = oracle.features(lap, eigval, eigvec, NU, KAPPA, x_is_train=False, edge_value=ev.to(dtype), edge_index=ei,
                            bump_scale=BUMP_SCALE, bump_decay=BUMP_DECAY)
    z_te = kernel.features(xt.to(DEV)).cpu()
    z_tr = kernel.features(x).cpu()
    cross, cross_o = z_te @ z_tr.T, zo_te @ zo_tr.T
    assert float((cross - cross_o).abs().max() / cross_o.abs().max()) < 1e-6
    # kernel forward: train x train is a low-rank root, test x train a matmul operator
    ktt = kernel(x, x)
    assert ktt.shape == (x.shape[0], x.shape[0])
    kst = kernel(xt.to(DEV), x).to_dense().cpu()
    assert float((kst - cross_o).abs().max() / cross_o.abs().max()) < 1e-6
    # posterior mean / variance at the held-out points (north_star: within 1e-4)
    mean_o, cov_o = oracle.low_rank_posterior(zo_tr, zo_te, dumbbell["train_y"].to(dtype), OUTPUTSCALE, NOISE)
    with torch.no_grad():
        model.posterior(xt.to(DEV))
        mean = model.posterior_mean.cpu()
        var = model.posterior_covar.to_dense().diagonal().cpu()
    assert float((mean - mean_o).abs().max() / mean_o.abs().max()) < 1e-4
    assert float((var - cov_o.diagonal()).abs().max() / cov_o.diagonal().abs().max()) < 1e-4


def test_posterior_fp32_within_tolerance(dumbbell):
    """fp32 is the reference's arithmetic.  Its own eval path is a DENSE fp32 eigh
    (riemann_kernel.py:124) whose 50 eigenvectors carry ~1e-3 of rounding, so the fp32 posterior cannot agree with the fp64
    truth to the north-star's 1e-4 -- for the reference itself either.  The test therefore runs the reference's dense path in
    fp32 beside ours (the oracle restatement evaluated in float32 on the CPU) and requires (a) our fp32 mean AND variance to
    be as close to the fp64 truth as the reference's own fp32 path is (within 3x of its error, or 1e-4 if that is larger),
    and (b) a hard ceiling of 1e-2 on the mean.  Measured on B200 (round 2): mean 3.2e-3 (reference's fp32 path 4.2e-3);
    the fp32 VARIANCE k** - k*^T (K + s I)^-1 k* cancels catastrophically at noise 1e-2 -- relative error 6.6 here and 6.3 for
    the reference's own fp32 path -- so it is only required to be no worse than the reference's."""
    model, kernel, x, y = _build(dumbbell, torch.float32)
    xt = dumbbell["test_x"]
    model.eval()
    model.likelihood.eval()

    def oracle_posterior(dtype):
        xo, lap, eigval, eigvec = _oracle_side(dumbbell, dtype)
        ev, ei = oracle.knn_search(xo.float(), xt.float(), K)
        zo_tr = oracle.features(lap, eigval, eigvec, NU, KAPPA, x_is_train=True)
        zo_te = oracle.features(lap, eigval, eigvec, NU, KAPPA, x_is_train=False, edge_value=ev.to(dtype), edge_index=ei,
                                bump_scale=BUMP_SCALE, bump_decay=BUMP_DECAY)
        m, c = oracle.low_rank_posterior(zo_tr, zo_te, dumbbell["train_y"].to(dtype), OUTPUTSCALE, NOISE)
        return m.double(), c.diagonal().double()

    mean_o, var_o = oracle_posterior(torch.float64)
    mean_r32, var_r32 = oracle_posterior(torch.float32)          # the reference's dense path in its own arithmetic
    with torch.no_grad():
        model.posterior(xt.float().to(DEV))
        mean = model.posterior_mean.cpu().double()
        var = model.posterior_covar.to_dense().diagonal().cpu().double()
    err = lambda a, b: float((a - b).abs().max() / b.abs().max())
    e_mean, e_var = err(mean, mean_o), err(var, var_o)
    r_mean, r_var = err(mean_r32, mean_o), err(var_r32, var_o)
    print(f"fp32 posterior vs fp64 truth: ours mean {e_mean:.2e} var {e_var:.2e}; reference fp32 dense path mean {r_mean:.2e} var {r_var:.2e}")
    assert e_mean < max(1e-4, 3 * r_mean) and e_mean < 1e-2
    assert e_var < max(1e-4, 3 * r_var)


def test_training_loss_matches_dense_reference_and_decreases(dumbbell):
    """manifold_informed_train (train_model.py:49-109): first loss value against the oracle's dense evaluation of
    1/2 [y^T Q y - log|Q| + n log 2pi] / n with Q = Noise(Scale(Matern precision)), then a few Adam steps."""
    import manifold_gp_b200 as mgp
    dtype = torch.float64
    model, kernel, x, y = _build(dumbbell, dtype)
    n = x.shape[0]
    xo, lap, _, _ = _oracle_side(dumbbell, dtype)
    yo = dumbbell["train_y"].to(dtype)

    def q_matmul(v):
        p = lambda t: oracle.scale_matmul(lambda u: oracle.precision_matmul(lap, NU, KAPPA, u), OUTPUTSCALE, t)
        return oracle.noise_matmul(p, NOISE, v)
    qd = oracle.dense_from_matmul(q_matmul, n, dtype)
    qd = 0.5 * (qd + qd.T)
    loss_o = 0.5 * (yo @ (qd @ yo) - torch.logdet(qd) + n * math.log(2 * math.pi)) / n
    model.train()
    with mgp.settings.max_cholesky_size(4000):
        prec = model.precision()
        loss = 0.5 * (torch.dot(y, prec.matmul(y.view(-1, 1)).squeeze()) - prec.inv_quad_logdet(logdet=True)[1]
                      + n * math.log(2 * math.pi)) / n
    assert abs(float(loss) - float(loss_o)) < 1e-6 * max(1.0, abs(float(loss_o)))
    # gradients reach all four hyper-parameters through the CUDA operators
    loss.backward()
    grads = {n_: p.grad for n_, p in model.named_parameters() if p.grad is not None}
    for name in ("covar_module.base_kernel.raw_graphbandwidth", "covar_module.base_kernel.raw_lengthscale",
                 "covar_module.raw_outputscale", "likelihood.noise_covar.raw_noise"):
        assert name in grads and torch.isfinite(grads[name]).all() and float(grads[name].abs().sum()) > 0, name
    from manifold_gp_b200.utils import manifold_informed_train
    opt = torch.optim.Adam(model.parameters(), lr=2e-2)
    l_start = manifold_informed_train(model, opt, max_iter=0, tolerance=0.0, max_cholesky=4000)
    l_end = manifold_informed_train(model, opt, max_iter=8, tolerance=0.0, max_cholesky=4000)
    assert math.isfinite(l_start) and math.isfinite(l_end) and l_end < l_start + 0.05
    # CG + stochastic Lanczos branch (max_cholesky below N): same objective within Monte-Carlo error
    frozen = torch.optim.Adam(model.parameters(), lr=0.0)
    with mgp.settings.num_trace_samples(64), mgp.settings.max_lanczos_quadrature_iterations(60):
        l_cg = manifold_informed_train(model, frozen, max_iter=0, tolerance=0.0, max_cholesky=100, cg_tolerance=1e-6, cg_max_iter=4000)
    l_ch = manifold_informed_train(model, frozen, max_iter=0, tolerance=0.0, max_cholesky=4000)
    assert abs(l_cg - l_ch) < 0.05 * max(1.0, abs(l_ch))


def test_graph_file_reuse(dumbbell, tmp_path):
    """graph_file: the second kernel loads the saved graph instead of searching and produces the same
    operator (edge list, values, row order -> identical precision matvec)."""
    import manifold_gp_b200 as mgp
    x = dumbbell["train_x"].double().to(DEV)
    path = str(tmp_path / "dumbbell_k10.pt")
    kw = dict(nu=NU, x=x, nearest_neighbors=K, laplacian_normalization="symmetric", num_modes=MODES, graph_file=path)
    k1 = mgp.RiemannMaternKernel(**kw).to(DEV).double()
    import os
    assert os.path.exists(path)
    k2 = mgp.RiemannMaternKernel(**kw).to(DEV).double()
    assert k2.knn.last_search is None, "the second kernel searched again instead of loading the graph file"
    assert torch.equal(k1.edge_index, k2.edge_index) and torch.equal(k1.edge_value, k2.edge_value)
    for kk in (k1, k2):
        kk.graphbandwidth = torch.tensor([[EPS]], dtype=torch.float64, device=DEV)
        kk.lengthscale = torch.tensor([[KAPPA]], dtype=torch.float64, device=DEV)
    v = torch.randn(x.shape[0], 3, dtype=torch.float64, device=DEV, generator=torch.Generator(device=DEV).manual_seed(0))
    assert torch.equal(k1.precision()._matmul(v), k2.precision()._matmul(v))
