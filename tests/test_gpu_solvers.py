"""GPU parity of the CUDA CG / SLQ / Lanczos drivers against the oracle's restatement of linear_operator's
algorithms and against the reference's dense ground truth (golden fixtures).  North-star tolerance: 1e-4 relative."""
import warnings

import pytest
import torch

import oracle
from conftest import gtag, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _setup(g, normalization="symmetric", eps=0.5, kappa=1.3, nu=2, dtype=torch.float64):
    import manifold_gp_b200 as mgp
    idx = torch.from_numpy(g["idx"]).long()
    val = torch.from_numpy(g["val"]).to(dtype)
    n = g["V"].shape[0]
    lap = mgp.GraphLaplacianOperator(val.to(DEV), idx.to(DEV), n, torch.tensor([[eps]], dtype=dtype, device=DEV), normalization)
    prec = mgp.PrecisionMaternOperator(lap, nu, torch.tensor([[kappa]], dtype=dtype, device=DEV))
    olap = oracle.LaplacianOracle(val, idx, n, torch.tensor(eps, dtype=dtype), normalization, True)
    oprec = lambda v: oracle.precision_matmul(olap, nu, kappa, v)
    return lap, prec, olap, oprec, n


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("normalization", ["symmetric", "randomwalk"])
def test_cg_vs_reference_dense_solve_and_oracle(golden_k10, normalization, dtype):
    from manifold_gp_b200 import solvers
    g = golden_k10
    lap, prec, olap, oprec, n = _setup(g, normalization, dtype=dtype)
    V = torch.from_numpy(g["V"]).to(dtype)
    tol = 1e-5
    x, info = solvers.linear_cg(prec, V.to(DEV), tolerance=tol, max_iter=1000, return_info=True)
    tag = gtag(0.5, 1.3, 2, normalization, True)
    assert info["converged"]
    assert rel_err(x, g[f"{tag}_Pinv_V"]) < 1e-4                       # vs the reference's dense solve
    ox, oinfo = oracle.linear_cg(oprec, V, tolerance=tol, max_iter=1000, return_info=True)
    assert rel_err(x, ox) < 1e-4                                       # vs the oracle's mBCG
    if dtype == torch.float64:
        assert info["iterations"] == oinfo["iterations"]               # same stopping rule, same iteration count
        assert rel_err(info["residual_norm"], oinfo["residual_norm"]) < 1e-6
    # 1-D rhs and the generic (un-fused) operator path through a wrapper
    import manifold_gp_b200 as mgp
    x1 = solvers.linear_cg(prec, V[:, 1].to(DEV), tolerance=tol, max_iter=1000)
    assert x1.shape == (n,) and rel_err(x1, g[f"{tag}_Pinv_V"][:, 1]) < 1e-4
    sc = mgp.ScaleWrapperOperator(prec, torch.tensor(2.0, dtype=dtype, device=DEV))
    xs = solvers.linear_cg(sc, V.to(DEV), tolerance=tol, max_iter=1000)
    assert rel_err(xs, g[f"{tag}_Pinv_V"] / 2.0) < 1e-4


def test_cg_stopping_rules_match_oracle(golden_k10):
    from manifold_gp_b200 import solvers
    g = golden_k10
    lap, prec, olap, oprec, n = _setup(g)
    gen = torch.Generator().manual_seed(0)
    B = torch.randn(n, 5, generator=gen, dtype=torch.float64)
    B[:, 2] = 0                                                        # zero right-hand side column
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for tol, max_iter in ((1e3, 100), (1e-30, 7), (1e-3, 1000), (1.0, 1000)):
            x, info = solvers.linear_cg(prec, B.to(DEV), tolerance=tol, max_iter=max_iter, max_tridiag_iter=min(20, max_iter), return_info=True)
            ox, oinfo = oracle.linear_cg(oprec, B, tolerance=tol, max_iter=max_iter, max_tridiag_iter=min(20, max_iter), return_info=True)
            assert info["iterations"] == oinfo["iterations"], (tol, max_iter)
            assert info["converged"] == oinfo["converged"]
            assert rel_err(x, ox) < 1e-8
            assert float(x[:, 2].abs().max()) == 0.0 and bool(torch.isfinite(x).all())


def test_cg_many_columns(golden_k10):
    """100 one-hot right-hand sides (the _average_variance call, precision_matern_operator.py:45-53) and > 128 columns."""
    from manifold_gp_b200 import solvers
    g = golden_k10
    lap, prec, olap, oprec, n = _setup(g)
    for c in (100, 130):
        B = torch.zeros(n, c, dtype=torch.float64)
        B[torch.arange(c) * 7, torch.arange(c)] = 1.0
        x = solvers.linear_cg(prec, B.to(DEV), tolerance=1e-5, max_iter=1000)
        ref = torch.linalg.solve(oracle.dense_from_matmul(oprec, n), B)
        assert rel_err(x, ref) < 1e-4


def test_cg_tridiag_and_slq_logdet_vs_oracle_and_dense(golden_k10):
    import manifold_gp_b200 as mgp
    from manifold_gp_b200 import solvers
    g = golden_k10
    lap, prec, olap, oprec, n = _setup(g)
    gen = torch.Generator().manual_seed(11)
    probes = torch.randn(n, 30, generator=gen, dtype=torch.float64)
    unit = probes / probes.norm(dim=0, keepdim=True)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        _, T = solvers.linear_cg(prec, unit.to(DEV), n_tridiag=30, tolerance=1e-5, max_iter=500, max_tridiag_iter=40)
        _, oT = oracle.linear_cg(oprec, unit, n_tridiag=30, tolerance=1e-5, max_iter=500, max_tridiag_iter=40)
    assert T.shape == oT.shape
    assert rel_err(T, oT) < 1e-6
    y = torch.from_numpy(g["V"][:, :1])
    with mgp.settings.max_cholesky_size(0), mgp.settings.cg_tolerance(1e-5), mgp.settings.max_cg_iterations(500), \
            mgp.settings.max_lanczos_quadrature_iterations(40):
        iq, ld = solvers.inv_quad_logdet(prec, inv_quad_rhs=y.to(DEV), logdet=True, probes=probes.to(DEV))
    oiq, old = oracle.inv_quad_logdet(oprec, n, inv_quad_rhs=y, logdet=True, probes=probes, tolerance=1e-5, max_iter=500,
                                      max_tridiag_iter=40, dtype=torch.float64)
    assert abs(float(ld) - float(old)) < 1e-6 * abs(float(old))        # same probes -> same estimate as the oracle
    assert abs(float(iq) - float(oiq)) < 1e-6 * abs(float(oiq))
    tag = gtag(0.5, 1.3, 2, "symmetric", True)
    ref_ld = float(g[f"{tag}_logdetP"])
    assert abs(float(ld) - ref_ld) / abs(ref_ld) < 0.03                # vs the reference's dense logdet (Monte-Carlo error)
    ref_iq = float((y * torch.from_numpy(g[f"{tag}_Pinv_V"][:, :1])).sum())
    assert abs(float(iq) - ref_iq) / abs(ref_iq) < 1e-4


def test_nll_gradients_through_cg_slq(golden_k10, dumbbell):
    """Precision-form NLL gradients on the CG/SLQ branch (size > max_cholesky_size) vs the reference's dense autograd:
    the yQy term is exact; the log-det term is a Hutchinson estimate (same estimator as linear_operator's)."""
    import math
    import manifold_gp_b200 as mgp
    g = golden_k10
    tag = gtag(0.5, 1.3, 2, "symmetric", True)
    dtype = torch.float64
    idx = torch.from_numpy(g["idx"]).long().to(DEV)
    val = torch.from_numpy(g["val"]).to(dtype).to(DEV)
    n = g["V"].shape[0]
    e = torch.tensor([[0.5]], dtype=dtype, device=DEV, requires_grad=True)
    kp = torch.tensor([[1.3]], dtype=dtype, device=DEV, requires_grad=True)
    oscale = torch.tensor(1.7, dtype=dtype, device=DEV, requires_grad=True)
    noise = torch.tensor(0.02, dtype=dtype, device=DEV, requires_grad=True)
    y = dumbbell["train_y"].double().to(DEV)
    torch.manual_seed(0)
    with mgp.settings.max_cholesky_size(0), mgp.settings.cg_tolerance(1e-5), mgp.settings.max_cg_iterations(2000), \
            mgp.settings.num_trace_samples(120), mgp.settings.max_lanczos_quadrature_iterations(60):
        lap = mgp.GraphLaplacianOperator(val, idx, n, e, "symmetric")
        prec = mgp.PrecisionMaternOperator(lap, 2, kp)
        op = mgp.NoiseWrapperOperator(mgp.ScaleWrapperOperator(prec, oscale, inverse_scale=True), noise)
        loss = 0.5 * sum([torch.dot(y, op.matmul(y.view(-1, 1)).squeeze()), -op.inv_quad_logdet(logdet=True)[1],
                          n * math.log(2 * math.pi)])
    grads = torch.autograd.grad(loss, [e, kp, oscale, noise])
    ref = float(g[f"{tag}_nll"])
    assert abs(loss.item() - ref) / abs(ref) < 0.02
    for a, b in zip(grads, g[f"{tag}_nll_grads"]):
        assert abs(a.item() - b) < 0.05 * max(1.0, abs(b)), (a.item(), b)


def test_lanczos_vs_oracle_and_reference_dense_eigh(golden_k10):
    import manifold_gp_b200 as mgp
    from manifold_gp_b200 import solvers
    g = golden_k10
    for dtype, otol in ((torch.float64, 1e-8), (torch.float32, 2e-4)):
        lap, prec, olap, oprec, n = _setup(g, dtype=dtype)
        gen = torch.Generator().manual_seed(5)
        v0 = torch.randn(n, generator=gen, dtype=dtype)
        q, t = solvers.lanczos_tridiag(lap, 120, init_vec=v0.to(DEV))
        j = q.shape[0]
        qd = q.double()
        assert float((qd @ qd.T - torch.eye(j, dtype=torch.float64, device=DEV)).abs().max()) < otol
        Aq = lap.matmul(q.T.contiguous()).double()
        assert float((qd @ Aq - t.double()).abs().max()) < otol * float(t.abs().max()) * 10
        if dtype == torch.float64:
            oq, ot = oracle.lanczos_tridiag(olap.matmul, 120, n, dtype=dtype, init_vec=v0)
            assert rel_err(torch.linalg.eigvalsh(t), torch.linalg.eigvalsh(ot)) < 1e-8
            assert rel_err(t.diagonal(), ot.diagonal()) < 1e-6
    # full-length Lanczos through the operator API: every low eigenpair of the reference's dense eigh
    lap, _, olap, _, n = _setup(g, dtype=torch.float64)
    tag = gtag(0.5, 1.3, 2, "symmetric", True)
    with mgp.settings.max_cholesky_size(0):
        evals, evecs = lap.diagonalization(method="lanczos", num_modes=None)
    ref = torch.from_numpy(g[f"{tag}_evals"])
    assert rel_err(evals[1:20], ref[1:20]) < 1e-4                     # north-star: top-m eigenvalues within 1e-4
    U = torch.from_numpy(g[f"{tag}_evecs"])
    for m in range(1, 6):
        a, b = evecs[:, m].cpu(), U[:, m] / U[:, m].norm()
        assert min(float((a - b).norm()), float((a + b).norm())) < 1e-4


@pytest.mark.parametrize("normalization", ["symmetric", "randomwalk"])
def test_diagonalization_api(golden_k10, normalization):
    """GraphLaplacianOperator.diagonalization(method, num_modes) -- graph_laplacian_operator.py:132-144."""
    import manifold_gp_b200 as mgp
    g = golden_k10
    lap, _, olap, _, n = _setup(g, normalization, dtype=torch.float64)
    tag = gtag(0.5, 1.3, 2, normalization, True)
    ref = torch.from_numpy(g[f"{tag}_evals"])
    with mgp.settings.max_cholesky_size(2000):                        # symeig branch (what test_laplacian.py:62 runs)
        evals, evecs = lap.diagonalization(num_modes=20)
    assert evals.shape == (20,) and evecs.shape == (n, 20) and float(evals[0]) == 0.0
    assert rel_err(evals[1:], ref[1:20]) < 1e-8
    with mgp.settings.max_cholesky_size(0):                           # lanczos branch, 3*num_modes steps
        evals_l, evecs_l = lap.diagonalization(num_modes=20)
    assert evals_l.shape == (20,) and evecs_l.shape == (n, 20)
    # with only 60 Lanczos steps the Ritz values are upper bounds that interlace the spectrum (extremal ones converge first)
    full = torch.linalg.eigvalsh(olap.symmetric_twin().dense())
    assert float(evals_l[1]) >= float(full[1]) - 1e-9 and float(evals_l[-1]) <= float(full[-1]) + 1e-9


def test_schur_inner_cg_consumer(golden_k10):
    import manifold_gp_b200 as mgp
    g = golden_k10
    _, prec, _, oprec, n = _setup(g, dtype=torch.float32)
    mask = torch.from_numpy(g["mask"]).to(DEV)
    V = torch.from_numpy(g["V"]).float().to(DEV)
    sch = mgp.SchurComplementOperator(prec, mask)
    with mgp.settings.max_cholesky_size(0), mgp.settings.cg_tolerance(1e-5), mgp.settings.max_cg_iterations(2000):
        out = sch.matmul(V[mask])
    tag = gtag(0.5, 1.3, 2, "symmetric", True)
    assert rel_err(out, g[f"{tag}_PschurV"]) < 1e-4


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_cg_cuda_graph_replay_is_bitwise_eager(dtype):
    """Long solves replay chunks of iterations from a CUDA graph; the arithmetic must be exactly the eager loop's
    (same kernels, same order), including the pending x update after the iteration that sets the done flag and the no-op
    replays past convergence.  Also: the split update (rupdate / pxupdate) against the oracle's mBCG."""
    import manifold_gp_b200 as mgp
    from manifold_gp_b200 import settings, solvers
    n, k = 30000, 12
    x = oracle.datasets.torus(n, seed=5)
    idx, val = mgp.NearestNeighbors(x.to(DEV)).graph(k)
    lap = mgp.GraphLaplacianOperator(val.to(dtype), idx, n, torch.tensor([[0.08]], dtype=dtype, device=DEV), "symmetric")
    prec = mgp.PrecisionMaternOperator(lap, 2, torch.tensor([[1.0]], dtype=dtype, device=DEV))
    g = torch.Generator().manual_seed(3)
    B = torch.randn(n, 16, generator=g).to(dtype).to(DEV)
    tol = 1e-5
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with settings.cg_cuda_graph(False):
            xe, ie = solvers.linear_cg(prec, B, tolerance=tol, max_iter=3000, return_info=True)
        with settings.cg_cuda_graph(True):
            xg, ig = solvers.linear_cg(prec, B, tolerance=tol, max_iter=3000, return_info=True)
            xg2, ig2 = solvers.linear_cg(prec, B, tolerance=tol, max_iter=3000, return_info=True)
    assert ie["iterations"] > 4 * settings.cg_check_interval.value(), "problem too easy to exercise the graph path"
    assert ie["iterations"] == ig["iterations"] == ig2["iterations"] and ie["converged"] and ig["converged"]
    assert torch.equal(xe, xg) and torch.equal(xg, xg2)
    res = (prec.matmul(xg) - B).double().norm(dim=0) / B.double().norm(dim=0)
    assert float(res.max()) < 5e-3
    # oracle mBCG on the same operator (fp64 only: iteration counts are then identical)
    if dtype == torch.float64:
        olap = oracle.LaplacianOracle(val.cpu().double(), idx.cpu(), n, torch.tensor(0.08, dtype=torch.float64), "symmetric", True)
        ox, oi = oracle.linear_cg(lambda t: oracle.precision_matmul(olap, 2, 1.0, t), B.cpu(), tolerance=tol, max_iter=3000,
                                  return_info=True)
        assert oi["iterations"] == ig["iterations"]
        assert rel_err(xg, ox) < 1e-6


@pytest.mark.parametrize("dtype,rtol,etol", [(torch.float64, 1e-9, 1e-7), (torch.float32, 2e-6, 2e-4)])
def test_smallest_eigenpairs_chebyshev_vs_dense_eigh(dtype, rtol, etol):
    """solvers.smallest_eigenpairs (what RiemannKernel.eval uses beyond dense_eigh_limit) against the reference's dense eigh
    (riemann_kernel.py:124) on a graph where 3m-step Lanczos cannot reach the bottom of the spectrum (N = 6000, m = 48)."""
    import manifold_gp_b200 as mgp
    from manifold_gp_b200 import solvers
    n, m = 6000, 48
    x = oracle.datasets.torus(n, seed=5).to(dtype).to(DEV)
    idx, val = mgp.NearestNeighbors(x.float()).graph(12)
    lap = mgp.GraphLaplacianOperator(val.to(dtype), idx, n, torch.tensor([[0.25]], dtype=dtype, device=DEV), "symmetric", True)
    dense = lap.to_dense()
    ev_ref, vec_ref = torch.linalg.eigh((0.5 * (dense + dense.T)).double())
    evals, evecs, info = solvers.smallest_eigenpairs(lap, m, rtol=rtol, return_info=True,
                                                     generator=torch.Generator(device=DEV).manual_seed(0))
    lam_max = float(ev_ref[-1])
    assert info["residual"] <= rtol * info["lam_max"], info
    assert float((evals.double() - ev_ref[:m]).abs().max()) < etol * lam_max
    # the invariant subspace matches (eigenvalues come in near-degenerate pairs on the torus: compare projectors)
    p = evecs.double() @ evecs.double().T
    p_ref = vec_ref[:, :m] @ vec_ref[:, :m].T
    gap_ok = float(ev_ref[m] - ev_ref[m - 1]) > 1e-3 * lam_max
    if gap_ok:
        assert float((p - p_ref).norm() / p_ref.norm()) < (1e-5 if dtype == torch.float64 else 5e-2)
    # and the operator-level entry point used by the kernel
    ev2, vec2 = lap.diagonalization(method="chebyshev", num_modes=m)
    assert ev2.shape == (m,) and vec2.shape == (n, m) and float(ev2[0]) == 0.0
    assert float((ev2[1:].double() - ev_ref[1:m]).abs().max()) < max(etol, 1e-5) * lam_max
