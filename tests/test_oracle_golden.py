"""Pin the oracle (oracle/operators.py, oracle/kernel.py) against the REFERENCE's dense operators
(test/_dense_operators.py) through the committed golden fixtures (tests/golden/make_golden.py).

Mirrors the reference's own test method (test/test_laplacian.py:58-68, test/_test_functions.py) with real
assertions: fp64 oracle vs fp64 dense reference at 1e-10 relative; fp32 oracle at 1e-5 relative.
"""
import math

import pytest
import torch

import oracle
from conftest import NORMALIZATIONS, PARAM_SETS_K10, gtag, rel_err

TOL = {torch.float64: 1e-10, torch.float32: 1e-5}


def _lap(g, eps, normalization, self_loops, dtype):
    idx = torch.from_numpy(g["idx"]).long()
    val = torch.from_numpy(g["val"]).to(dtype)
    n = g["V"].shape[0]
    return oracle.LaplacianOracle(val, idx, n, torch.tensor(eps, dtype=dtype), normalization, self_loops)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("self_loops", [True, False])
@pytest.mark.parametrize("normalization", NORMALIZATIONS)
@pytest.mark.parametrize("eps,kappa,nu", PARAM_SETS_K10)
def test_laplacian_and_precision_vs_reference_dense(golden_k10, eps, kappa, nu, normalization, self_loops, dtype):
    g = golden_k10
    tag = gtag(eps, kappa, nu, normalization, self_loops)
    tol = TOL[dtype]
    lap = _lap(g, eps, normalization, self_loops, dtype)
    V = torch.from_numpy(g["V"]).to(dtype)
    assert rel_err(lap.degree_unnorm_mat, g[f"{tag}_deg_unnorm"]) < tol
    assert rel_err(lap.degree_mat, g[f"{tag}_deg"]) < tol
    assert rel_err(lap.matmul(V[:, 1:]), g[f"{tag}_LV"][:, 1:]) < tol                     # test_mv
    assert rel_err(lap.transpose().matmul(V[:, 1:]), g[f"{tag}_LtV"][:, 1:]) < tol        # test_mv_transpose
    # 1-D rhs = train_y, a smooth function: L y cancels ~100x (|L y| << |diag*y|), which amplifies fp32 rounding in ANY
    # summation order -- measure that column against the size of the terms being summed (backward-error sense)
    y1 = lap.matmul(V[:, 0])
    assert y1.shape == (lap.n,)
    scale = float((lap.laplacian_diag * V[:, 0]).double().norm())
    assert float((y1.double() - torch.from_numpy(g[f"{tag}_LV"][:, 0])).norm()) / scale < tol
    # diag: for randomwalk the dense diagonal equals the symmetric-form diagonal (similarity transform)
    assert rel_err(lap.laplacian_diag, g[f"{tag}_Ldiag"]) < tol            # test_diag
    P = lambda v: oracle.precision_matmul(lap, nu, kappa, v)
    ptol = tol * (10 if dtype == torch.float32 and nu == 3 else 1)          # nu=3, eps=0.05: cond ~1e9 amplifies fp32 rounding
    assert rel_err(P(V), g[f"{tag}_PV"]) < ptol
    oscale, noise = 1.7, 0.02
    Pdiv = lambda v: oracle.scale_matmul(P, oscale, v, inverse_scale=True)
    Pmul = lambda v: oracle.scale_matmul(P, oscale, v)
    assert rel_err(Pdiv(V), g[f"{tag}_PdivV"]) < ptol
    assert rel_err(Pmul(V), g[f"{tag}_PmulV"]) < ptol
    if dtype == torch.float64 or nu < 3:
        assert rel_err(oracle.noise_matmul(Pdiv, noise, V), g[f"{tag}_PnoisyV"]) < ptol * 10


@pytest.mark.parametrize("normalization", NORMALIZATIONS)
def test_schur_vs_reference_dense(golden_k10, normalization):
    g = golden_k10
    eps, kappa, nu = 0.5, 1.3, 2
    tag = gtag(eps, kappa, nu, normalization, True)
    lap = _lap(g, eps, normalization, True, torch.float64)
    V = torch.from_numpy(g["V"])
    mask = torch.from_numpy(g["mask"])
    P = lambda v: oracle.precision_matmul(lap, nu, kappa, v)
    out = oracle.schur_matmul(P, mask, V[mask])
    assert rel_err(out, g[f"{tag}_PschurV"]) < 1e-9


@pytest.mark.parametrize("self_loops", [True, False])
@pytest.mark.parametrize("normalization", NORMALIZATIONS)
def test_grad_eps_vs_reference_dense(golden_k10, normalization, self_loops):
    """test_grad (test/_test_functions.py:59-74): d/d eps of sum(L^T v) by autograd through the oracle."""
    g = golden_k10
    eps_v, kappa, nu = 0.5, 1.3, 2
    tag = gtag(eps_v, kappa, nu, normalization, self_loops)
    eps = torch.tensor(eps_v, dtype=torch.float64, requires_grad=True)
    idx = torch.from_numpy(g["idx"]).long()
    val = torch.from_numpy(g["val"]).double()
    lap = oracle.LaplacianOracle(val, idx, g["V"].shape[0], eps, normalization, self_loops)
    v = torch.from_numpy(g["V"])[:, :1]
    (ge,) = torch.autograd.grad(lap.transpose().matmul(v).sum(), eps)
    ref = float(g[f"{tag}_grad_eps_sumLtv"].reshape(-1)[0])
    # sum(L^T v) is ~0 for the random-walk operator (rows of L sum to 0): compare absolutely there
    assert abs(float(ge) - ref) < 1e-9 * max(1.0, abs(ref))


@pytest.mark.parametrize("normalization", NORMALIZATIONS)
def test_nll_and_grads_vs_reference_dense(golden_k10, dumbbell, normalization):
    """test_ml (test/_test_functions.py:77-81): precision-form NLL through Scale(inverse)+Noise wrappers, fp64,
    with the dense-Cholesky branch of inv_quad_logdet (what the reference's notebooks run: max_cholesky >= N)."""
    g = golden_k10
    eps_v, kappa_v, nu = 0.5, 1.3, 2
    tag = gtag(eps_v, kappa_v, nu, normalization, True)
    eps = torch.tensor(eps_v, dtype=torch.float64, requires_grad=True)
    kappa = torch.tensor(kappa_v, dtype=torch.float64, requires_grad=True)
    oscale = torch.tensor(1.7, dtype=torch.float64, requires_grad=True)
    noise = torch.tensor(0.02, dtype=torch.float64, requires_grad=True)
    idx = torch.from_numpy(g["idx"]).long()
    val = torch.from_numpy(g["val"]).double()
    n = g["V"].shape[0]
    lap = oracle.LaplacianOracle(val, idx, n, eps, normalization, True)
    P = lambda v: oracle.precision_matmul(lap, nu, kappa, v)
    Pd = lambda v: oracle.scale_matmul(P, oscale, v, inverse_scale=True)
    Pn = lambda v: oracle.noise_matmul(Pd, noise, v)
    y = dumbbell["train_y"].double()
    dense = oracle.dense_from_matmul(Pn, n)
    loss = 0.5 * sum([torch.dot(y, Pn(y)), -torch.logdet(dense), n * math.log(2 * math.pi)])
    grads = torch.autograd.grad(loss, [eps, kappa, oscale, noise])
    assert abs(loss.item() - float(g[f"{tag}_nll"])) < 1e-8 * abs(float(g[f"{tag}_nll"]))
    for a, b in zip(grads, g[f"{tag}_nll_grads"]):
        assert abs(a.item() - b) < 1e-7 * max(1.0, abs(b))


@pytest.mark.parametrize("normalization", NORMALIZATIONS)
def test_eigen_and_out_of_sample_vs_reference_dense(golden_k10, normalization):
    """test_eigen + test_outofsample (test/_test_functions.py:107-164)."""
    g = golden_k10
    eps, kappa, nu = 0.5, 1.3, 2
    tag = gtag(eps, kappa, nu, normalization, True)
    lap = _lap(g, eps, normalization, True, torch.float64)
    sym = lap.symmetric_twin()
    w, U = torch.linalg.eigh(sym.dense())
    m = 20
    assert rel_err(w[1:m], g[f"{tag}_evals"][1:m]) < 1e-8      # similar matrices share the spectrum
    if normalization == "randomwalk":
        U = U * lap.degree_mat.pow(-0.5).view(-1, 1)
        U = torch.nn.functional.normalize(U, p=2, dim=0)
    ref_U = torch.from_numpy(g[f"{tag}_evecs"])
    for j in range(1, 6):   # non-degenerate low modes, up to sign (and scale for the non-normalised eig output)
        a, b = U[:, j], ref_U[:, j] / ref_U[:, j].norm()
        assert min((a - b).norm(), (a + b).norm()) < 1e-6
    # out-of-sample: the extension matrix applied to the reference's eigenvectors
    ev = torch.from_numpy(g["oos_edge_value"]).double()
    ei = torch.from_numpy(g["oos_edge_index"]).long()
    Uref = torch.zeros(lap.n, 6, dtype=torch.float64)
    Uref[:, :6] = ref_U
    ext = lap.out_of_sample(Uref, ev, ei)
    assert rel_err(ext, g[f"{tag}_ext_evecs"][:, :6]) < 1e-9


def test_k50_test_laplacian_configuration(golden_k50):
    """The exact configuration of test/test_laplacian.py:34-50: k=50, nu=1, eps=0.5, kappa=0.5, self_loops=False."""
    g = golden_k50
    for normalization in NORMALIZATIONS:
        tag = gtag(0.5, 0.5, 1, normalization, False)
        for dtype in (torch.float64, torch.float32):
            lap = _lap(g, 0.5, normalization, False, dtype)
            V = torch.from_numpy(g["V"]).to(dtype)
            assert rel_err(lap.matmul(V), g[f"{tag}_LV"]) < TOL[dtype]
            assert rel_err(lap.transpose().matmul(V), g[f"{tag}_LtV"]) < TOL[dtype]
            assert rel_err(oracle.precision_matmul(lap, 1, 0.5, V), g[f"{tag}_PV"]) < TOL[dtype]


def test_golden_graph_matches_oracle_graph(golden_k10, dumbbell):
    idx, val = oracle.knn_graph(dumbbell["train_x"], 10)
    assert torch.equal(idx, torch.from_numpy(golden_k10["idx"]).long())
    assert torch.equal(val, torch.from_numpy(golden_k10["val"]))
    # upper triangular, lexicographically sorted, unique
    key = idx[0] * dumbbell["train_x"].shape[0] + idx[1]
    assert bool((idx[0] < idx[1]).all()) and bool((key[1:] > key[:-1]).all())
