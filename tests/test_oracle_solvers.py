"""Pin oracle CG / Lanczos / SLQ against dense linear algebra on the reference's fixture (dumbbell, k=10).
linear_operator is third party and absent (parity unpinned vs the reference) -- the dense answers
are the reference's own ground truth in test_solve / test_eigen / test_ml (test/_test_functions.py:47,107,77)."""
import warnings

import pytest
import torch

import oracle
from conftest import gtag, rel_err


def _setup(g, normalization="symmetric", eps=0.5, kappa=1.3, nu=2, dtype=torch.float64):
    idx = torch.from_numpy(g["idx"]).long()
    val = torch.from_numpy(g["val"]).to(dtype)
    n = g["V"].shape[0]
    lap = oracle.LaplacianOracle(val, idx, n, torch.tensor(eps, dtype=dtype), normalization, True)
    P = lambda v: oracle.precision_matmul(lap, nu, kappa, v)
    return lap, P, n


@pytest.mark.parametrize("normalization", ["symmetric", "randomwalk"])
def test_cg_solution_vs_reference_dense_solve(golden_k10, normalization):
    g = golden_k10
    lap, P, n = _setup(g, normalization)
    V = torch.from_numpy(g["V"])
    # NB the published mBCG zeroes alpha when p^T A p < eps=1e-10, so with unit-normalised right-hand sides the
    # residual cannot be driven much below ~1e-5..1e-6 here; 1e-5 is comfortably reachable.
    x, info = oracle.linear_cg(P, V, tolerance=1e-5, max_iter=4000, return_info=True)
    assert info["converged"] and info["iterations"] < 200
    tag = gtag(0.5, 1.3, 2, normalization, True)
    assert rel_err(x, g[f"{tag}_Pinv_V"]) < 1e-4           # north-star tolerance for CG solutions


def test_cg_stopping_rules():
    torch.manual_seed(0)
    A = torch.randn(60, 60, dtype=torch.float64)
    A = A @ A.T + 60 * torch.eye(60, dtype=torch.float64)
    b = torch.randn(60, 3, dtype=torch.float64)
    # at least min(10, max_iter-1)+1 iterations even if the tolerance is met immediately
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        _, info = oracle.linear_cg(lambda v: A @ v, b, tolerance=1e3, max_iter=100, return_info=True)
        assert info["iterations"] == 11
        _, info = oracle.linear_cg(lambda v: A @ v, b, tolerance=1e-30, max_iter=7, max_tridiag_iter=7, return_info=True)
        assert info["iterations"] == 7 and not info["converged"]
    # zero rhs column -> zero solution, no NaN
    b0 = b.clone()
    b0[:, 1] = 0
    x = oracle.linear_cg(lambda v: A @ v, b0, tolerance=1e-5, max_iter=200)
    assert torch.isfinite(x).all() and float(x[:, 1].abs().max()) == 0.0
    assert rel_err(x, torch.linalg.solve(A, b0)) < 1e-4
    # 1-D rhs
    x1 = oracle.linear_cg(lambda v: A @ v, b[:, 0], tolerance=1e-5, max_iter=200)
    assert x1.shape == (60,)


def test_cg_tridiag_reproduces_lanczos_spectrum():
    """mBCG tridiagonals: eig(T) after n steps on a small SPD matrix equals eig(A) (exact Lanczos property)."""
    torch.manual_seed(1)
    n = 12
    A = torch.randn(n, n, dtype=torch.float64)
    A = A @ A.T + torch.eye(n, dtype=torch.float64)
    b = torch.randn(n, 2, dtype=torch.float64)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        _, T = oracle.linear_cg(lambda v: A @ v, b, n_tridiag=2, tolerance=1e-30, max_iter=n, max_tridiag_iter=n)
    assert T.shape == (2, n, n)
    for p in range(2):
        assert torch.allclose(torch.linalg.eigvalsh(T[p]), torch.linalg.eigvalsh(A), rtol=1e-6)


def test_slq_logdet_vs_reference_dense_logdet(golden_k10):
    g = golden_k10
    lap, P, n = _setup(g, "symmetric")
    gen = torch.Generator().manual_seed(11)
    probes = torch.randn(n, 30, generator=gen, dtype=torch.float64)
    _, ld = oracle.inv_quad_logdet(P, n, logdet=True, probes=probes, tolerance=1e-5, max_iter=500,
                                   max_tridiag_iter=60, dtype=torch.float64)
    ref = float(g[gtag(0.5, 1.3, 2, "symmetric", True) + "_logdetP"])
    assert abs(float(ld) - ref) / abs(ref) < 0.03      # Monte-Carlo error of 30 Hutchinson probes
    y = torch.from_numpy(g["V"][:, :1])
    iq, _ = oracle.inv_quad_logdet(P, n, inv_quad_rhs=y, logdet=False, tolerance=1e-5, max_iter=500)
    ref_iq = float((y * torch.from_numpy(g[gtag(0.5, 1.3, 2, "symmetric", True) + "_Pinv_V"][:, :1])).sum())
    assert abs(float(iq) - ref_iq) / abs(ref_iq) < 1e-4


def test_lanczos_vs_reference_dense_eigh(golden_k10):
    g = golden_k10
    lap, _, n = _setup(g, "symmetric")
    gen = torch.Generator().manual_seed(5)
    q, t = oracle.lanczos_tridiag(lap.matmul, 120, n, dtype=torch.float64, generator=gen)
    # orthonormal basis and T = Q^T A Q
    assert float((q.T @ q - torch.eye(q.shape[1], dtype=torch.float64)).abs().max()) < 1e-8
    assert float((q.T @ lap.matmul(q) - t).abs().max()) < 1e-6 * float(t.abs().max())
    # Ritz values: the extremal (largest) ones converge first; compare with the reference's dense spectrum
    w = torch.linalg.eigvalsh(lap.dense())
    ritz = torch.linalg.eigvalsh(t)
    assert torch.allclose(ritz[-2:], w[-2:], rtol=1e-5)
    # full-length Lanczos on a small operator with a well-separated spectrum recovers every eigenpair
    torch.manual_seed(3)
    sub = 30
    Qr, _ = torch.linalg.qr(torch.randn(sub, sub, dtype=torch.float64))
    wa = torch.linspace(0.0, 5.0, sub, dtype=torch.float64)
    A = (Qr * wa) @ Qr.T
    evals, evecs = oracle.lanczos_diagonalization(lambda v: A @ v, sub, sub, dtype=torch.float64, generator=gen)
    assert evals.shape[0] == sub
    assert torch.allclose(evals[1:], wa[1:], rtol=1e-6, atol=1e-8)
    assert float((A @ evecs[:, 1:] - evecs[:, 1:] * evals[1:]).abs().max()) < 1e-6


def test_zero_padded_block_with_rescaled_tolerance_is_the_same_solve():
    """The identity manifold_gp_b200.solvers.linear_cg relies on when it pads an 11-column block to 16 columns so that the
    128-bit SpMM kernels apply: zero right-hand sides have residual exactly 0, so the published mean-over-columns stopping
    rule on the padded block with tolerance * c / c_pad makes the same decisions as on the original block."""
    torch.manual_seed(1)
    n, c, cpad = 80, 11, 16
    A = torch.randn(n, n, dtype=torch.float64)
    A = A @ A.T + n * torch.eye(n, dtype=torch.float64)
    b = torch.randn(n, c, dtype=torch.float64)
    bp = torch.zeros(n, cpad, dtype=torch.float64)
    bp[:, :c] = b
    mm = lambda v: A @ v
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for tol in (1e-2, 1e-5, 1e-8):
            x, info = oracle.linear_cg(mm, b, tolerance=tol, max_iter=200, return_info=True)
            xp, infop = oracle.linear_cg(mm, bp, tolerance=tol * c / cpad, max_iter=200, return_info=True)
            assert info["iterations"] == infop["iterations"], tol
            assert torch.equal(xp[:, c:], torch.zeros(n, cpad - c, dtype=torch.float64))
            assert rel_err(xp[:, :c], x) < 1e-13
