"""GPU parity of the operators (through the C ABI) against the reference's dense operators (golden fixtures) and
the oracle.  Tolerances are the north-star's: matvecs 1e-5 relative in fp32, 1e-10 in fp64."""
import pytest
import torch

import oracle
from conftest import NORMALIZATIONS, PARAM_SETS_K10, gtag, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = {torch.float64: 1e-10, torch.float32: 1e-5}


def _ops(g, eps, kappa, nu, normalization, self_loops, dtype, requires_grad=False):
    import manifold_gp_b200 as mgp
    idx = torch.from_numpy(g["idx"]).long().to(DEV)
    val = torch.from_numpy(g["val"]).to(dtype).to(DEV)
    n = g["V"].shape[0]
    e = torch.tensor([[eps]], dtype=dtype, device=DEV, requires_grad=requires_grad)
    kp = torch.tensor([[kappa]], dtype=dtype, device=DEV, requires_grad=requires_grad)
    lap = mgp.GraphLaplacianOperator(val, idx, n, e, normalization, self_loops)
    prec = mgp.PrecisionMaternOperator(lap, nu, kp)
    return lap, prec, e, kp


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("self_loops", [True, False])
@pytest.mark.parametrize("normalization", NORMALIZATIONS)
@pytest.mark.parametrize("eps,kappa,nu", PARAM_SETS_K10)
def test_operators_vs_reference_dense(golden_k10, eps, kappa, nu, normalization, self_loops, dtype):
    import manifold_gp_b200 as mgp
    g = golden_k10
    tag = gtag(eps, kappa, nu, normalization, self_loops)
    tol = TOL[dtype]
    lap, prec, _, _ = _ops(g, eps, kappa, nu, normalization, self_loops, dtype)
    V = torch.from_numpy(g["V"]).to(dtype).to(DEV)
    assert lap.shape == (V.shape[0], V.shape[0])
    assert rel_err(lap.degree_unnorm_mat, g[f"{tag}_deg_unnorm"]) < tol
    assert rel_err(lap.degree_mat, g[f"{tag}_deg"]) < tol
    assert rel_err(lap.diagonal(), g[f"{tag}_Ldiag"]) < tol                                   # test_diag
    assert rel_err(lap.matmul(V[:, 1:]), g[f"{tag}_LV"][:, 1:]) < tol                          # test_mv
    assert rel_err(lap.T.matmul(V[:, 1:]), g[f"{tag}_LtV"][:, 1:]) < tol                       # test_mv_transpose
    assert rel_err(lap.matmul(V), g[f"{tag}_LV"]) < tol * 20                                   # incl. the smooth column
    y1 = lap.matmul(V[:, 0])                                                                   # 1-D rhs
    assert y1.shape == (V.shape[0],)
    scale = float((lap.laplacian_diag * V[:, 0]).double().norm())
    assert float((y1.double().cpu() - torch.from_numpy(g[f"{tag}_LV"][:, 0])).norm()) / scale < tol
    ptol = tol * (10 if dtype == torch.float32 and nu == 3 else 1)
    assert rel_err(prec.matmul(V), g[f"{tag}_PV"]) < ptol
    assert prec.T is prec
    oscale = torch.tensor(1.7, dtype=dtype, device=DEV)
    noise = torch.tensor(0.02, dtype=dtype, device=DEV)
    assert rel_err(mgp.ScaleWrapperOperator(prec, oscale).matmul(V), g[f"{tag}_PmulV"]) < ptol
    pdiv = mgp.ScaleWrapperOperator(prec, oscale, inverse_scale=True)
    assert rel_err(pdiv.matmul(V), g[f"{tag}_PdivV"]) < ptol
    if dtype == torch.float64 or nu < 3:
        assert rel_err(mgp.NoiseWrapperOperator(pdiv, noise).matmul(V), g[f"{tag}_PnoisyV"]) < ptol * 10
    # per-edge arrays the reference exposes
    olap = oracle.LaplacianOracle(torch.from_numpy(g["val"]).to(dtype), torch.from_numpy(g["idx"]).long(), V.shape[0],
                                  eps, normalization, self_loops)
    assert rel_err(lap.laplacian_triu, olap.laplacian_triu) < tol
    assert rel_err(lap.adjacency_mat, olap.adjacency_mat) < tol


@pytest.mark.parametrize("normalization", NORMALIZATIONS)
def test_k50_test_laplacian_configuration(golden_k50, normalization):
    """test/test_laplacian.py:34-50: k=50, nu=1, eps=0.5, kappa=0.5, self_loops=False."""
    g = golden_k50
    tag = gtag(0.5, 0.5, 1, normalization, False)
    for dtype in (torch.float64, torch.float32):
        lap, prec, _, _ = _ops(g, 0.5, 0.5, 1, normalization, False, dtype)
        V = torch.from_numpy(g["V"]).to(dtype).to(DEV)
        assert rel_err(lap.matmul(V[:, 1:]), g[f"{tag}_LV"][:, 1:]) < TOL[dtype]
        assert rel_err(lap.T.matmul(V[:, 1:]), g[f"{tag}_LtV"][:, 1:]) < TOL[dtype]
        assert rel_err(prec.matmul(V), g[f"{tag}_PV"]) < TOL[dtype]


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("ncols", [1, 2, 3, 4, 7, 8, 10, 16, 20, 33, 100])
def test_spmm_column_counts_vs_oracle(ncols, dtype):
    """Every kernel instantiation (vector widths 16/8/4, scalar widths 1..32, multi-pass) on a 30k-point torus."""
    import manifold_gp_b200 as mgp
    x = oracle.datasets.torus(30000, seed=1)
    oidx, oval = oracle.knn_graph(x, 16)
    n = x.shape[0]
    eps, kappa, nu = 0.12, 0.5, 2
    gen = torch.Generator().manual_seed(ncols)
    V = torch.randn(n, ncols, generator=gen, dtype=dtype)
    for normalization in NORMALIZATIONS:
        olap = oracle.LaplacianOracle(oval.double(), oidx, n, eps, normalization, True)
        ref_l = olap.matmul(V.double())
        ref_p = oracle.precision_matmul(olap, nu, kappa, V.double())
        lap = mgp.GraphLaplacianOperator(oval.to(dtype).to(DEV), oidx.to(DEV), n,
                                         torch.tensor([[eps]], dtype=dtype, device=DEV), normalization)
        prec = mgp.PrecisionMaternOperator(lap, nu, torch.tensor([[kappa]], dtype=dtype, device=DEV))
        assert rel_err(lap.matmul(V.to(DEV)), ref_l) < TOL[dtype]
        assert rel_err(prec.matmul(V.to(DEV)), ref_p) < TOL[dtype]
        # non-contiguous rhs (a column slice of a wider buffer)
        wide = torch.zeros(n, ncols + 3, dtype=dtype, device=DEV)
        wide[:, 1:1 + ncols] = V.to(DEV)
        assert rel_err(lap.matmul(wide[:, 1:1 + ncols]), ref_l) < TOL[dtype]


@pytest.mark.parametrize("n,k", [(120_000, 16), (1_000_000, 32)])
def test_dominant_kernel_vs_oracle_on_a_searched_graph(n, k):
    """The kernel the headline solve launches (lap_spmm_wi_kernel, paired-row walk, C = 16, selected by `auto` only on graphs that carry the
    Morton hint of NearestNeighbors.graph) DIRECTLY against the oracle's restatement of graph_laplacian_operator.py:108-124 /
    precision_matern_operator.py:26-37 on the same edge list, in fp32 (1e-5) and fp64 (1e-10), at a mid size and at the full
    cfg-C size (N = 1M, k = 32); plus the fused dot epilogue and the single-column tile SpMV on the same graph."""
    import manifold_gp_b200 as mgp
    from manifold_gp_b200 import graph
    x = oracle.datasets.torus(n, seed=0)
    idx, val = mgp.NearestNeighbors(x.to(DEV)).graph(k)
    eps = 0.0274 * (1_000_000 / n) ** 0.5
    kappa, nu = 0.5, 2
    olap = oracle.LaplacianOracle(val.double().cpu(), idx.cpu(), n, eps, "symmetric", True)
    V = torch.randn(n, 16, generator=torch.Generator().manual_seed(7), dtype=torch.float64)
    ref_l = olap.matmul(V)
    ref_p = oracle.precision_matmul(olap, nu, kappa, V)
    for dtype in (torch.float32, torch.float64):
        lap = mgp.GraphLaplacianOperator(val.to(dtype), idx, n, torch.tensor([[eps]], dtype=dtype, device=DEV), "symmetric", True)
        prec = mgp.PrecisionMaternOperator(lap, nu, torch.tensor([[kappa]], dtype=dtype, device=DEV))
        st = lap.structure
        assert st.perm is not None and st.tiles is not None and "wptr" in st.tiles
        # fp32: `auto` takes the paired-row walk of the kernel (the one the headline solve launches); the single-row walk of the
        # same kernel (all there is in fp64) is forced beside it
        expect = "lap_spmm_wi_kernel<pair>" if dtype == torch.float32 else "lap_spmm_wi_kernel"
        y = lap.matmul(V.to(dtype).to(DEV))
        assert graph.LAST_SPMM_KERNEL == expect
        assert rel_err(y, ref_l) < TOL[dtype]
        yp = prec.matmul(V.to(dtype).to(DEV))
        assert graph.LAST_SPMM_KERNEL == expect
        assert rel_err(yp, ref_p) < TOL[dtype]
        if dtype == torch.float32:
            graph.SPMM_KERNEL = "wi"
            try:
                yp1 = prec.matmul(V.to(dtype).to(DEV))
                assert graph.LAST_SPMM_KERNEL == "lap_spmm_wi_kernel"
            finally:
                graph.SPMM_KERNEL = "auto"
            assert rel_err(yp1, ref_p) < TOL[dtype]
        # solver-driver interface (internal order, caller-owned buffers, fused dot product) -- what CG actually calls
        xi = st.to_internal(V.to(dtype).to(DEV)).contiguous()
        out, tmp = torch.empty_like(xi), torch.empty_like(xi)
        dot = torch.zeros(16, dtype=dtype, device=DEV)
        prec._mgp_matvec(xi, out, tmp, dot_with=xi, dot_out=dot)
        assert rel_err(st.to_external(out), ref_p) < TOL[dtype]
        assert rel_err(dot, (V * ref_p).sum(0)) < TOL[dtype] * 10
        y1 = lap.matmul(V[:, :1].to(dtype).to(DEV))
        assert graph.LAST_SPMM_KERNEL == "lap_spmv_tile_kernel"
        assert rel_err(y1, ref_l[:, :1]) < TOL[dtype]
        del lap, prec, st, y, yp, xi, out, tmp
        torch.cuda.empty_cache()


def test_linearity_and_symmetry_at_scale():
    """Size-independent properties at N = 1M, k = 32 (BASELINE cfg-C): <u, P v> == <P u, v>, P(au+bv) == aPu + bPv."""
    import manifold_gp_b200 as mgp
    x = oracle.datasets.torus(1_000_000, seed=0).to(DEV)
    idx, val = mgp.NearestNeighbors(x).graph(32)
    n = x.shape[0]
    assert 16.5 < idx.shape[1] / n < 16.9                # M/N = 16.72 at N=1M, k=32 (85% of edges mutual)
    lap = mgp.GraphLaplacianOperator(val, idx, n, torch.tensor([[0.0274]], device=DEV), "symmetric")
    prec = mgp.PrecisionMaternOperator(lap, 2, torch.tensor([[0.5]], device=DEV))
    g = torch.Generator(device=DEV).manual_seed(3)
    u = torch.randn(n, 16, device=DEV, generator=g)
    v = torch.randn(n, 16, device=DEV, generator=g)
    pu, pv = prec.matmul(u), prec.matmul(v)
    a = (u.double() * pv.double()).sum()
    b = (pu.double() * v.double()).sum()
    assert abs(float(a - b)) / abs(float(a)) < 1e-5
    comb = prec.matmul(0.3 * u - 1.7 * v)
    assert rel_err(comb, 0.3 * pu - 1.7 * pv) < 1e-5
    # the Laplacian annihilates D^{1/2} 1 (constant function on the random-walk side)
    one = lap.degree_mat.sqrt().unsqueeze(-1)
    assert float(lap.matmul(one).norm() / (lap.laplacian_diag.unsqueeze(-1) * one).norm()) < 1e-5


def test_dot_epilogue_matches_separate_reduction():
    import manifold_gp_b200 as mgp
    from manifold_gp_b200 import graph
    x = oracle.datasets.torus(50000, seed=2)
    oidx, oval = oracle.knn_graph(x, 12)
    n = x.shape[0]
    for dtype in (torch.float32, torch.float64):
        lap = mgp.GraphLaplacianOperator(oval.to(dtype).to(DEV), oidx.to(DEV), n,
                                         torch.tensor([[0.1]], dtype=dtype, device=DEV), "symmetric")
        _, _, diag, a = lap._values()
        for c in (1, 5, 16, 32):
            X = torch.randn(n, c, dtype=dtype, device=DEV)
            dot = torch.zeros(c, dtype=dtype, device=DEV)
            Y = graph.lap_spmm(lap.structure, a, diag, X, dot_with=X, dot_out=dot)
            ref = (X.double() * Y.double()).sum(0)
            assert rel_err(dot, ref) < (1e-5 if dtype == torch.float32 else 1e-12)
            dot2 = torch.zeros(c, dtype=dtype, device=DEV)
            graph.lap_spmm(lap.structure, a, diag, X, dot_with=X, dot_out=dot2)
            assert torch.equal(dot, dot2), "dot epilogue must be deterministic run to run"


@pytest.mark.parametrize("self_loops", [True, False])
@pytest.mark.parametrize("normalization", NORMALIZATIONS)
def test_grad_eps_vs_reference_dense(golden_k10, normalization, self_loops):
    """test_grad (test/_test_functions.py:59-74): d/d eps of sum(L^T v) through the CUDA backward kernels."""
    g = golden_k10
    tag = gtag(0.5, 1.3, 2, normalization, self_loops)
    lap, _, e, _ = _ops(g, 0.5, 1.3, 2, normalization, self_loops, torch.float64, requires_grad=True)
    v = torch.from_numpy(g["V"])[:, :1].to(DEV)
    loss = lap.T.matmul(v).sum()
    (ge,) = torch.autograd.grad(loss, e)
    ref = float(g[f"{tag}_grad_eps_sumLtv"].reshape(-1)[0])
    assert abs(float(ge) - ref) < 1e-8 * max(1.0, abs(ref))


@pytest.mark.parametrize("normalization", NORMALIZATIONS)
def test_nll_and_grads_vs_reference_dense(golden_k10, dumbbell, normalization):
    """test_ml (test/_test_functions.py:77-104), Cholesky branch (max_cholesky >= N as in the notebooks), fp64:
    loss and d/d(eps, kappa, outputscale, noise) against the reference's dense autograd."""
    import math
    import manifold_gp_b200 as mgp
    g = golden_k10
    tag = gtag(0.5, 1.3, 2, normalization, True)
    lap, prec, e, kp = _ops(g, 0.5, 1.3, 2, normalization, True, torch.float64, requires_grad=True)
    oscale = torch.tensor(1.7, dtype=torch.float64, device=DEV, requires_grad=True)
    noise = torch.tensor(0.02, dtype=torch.float64, device=DEV, requires_grad=True)
    op = mgp.NoiseWrapperOperator(mgp.ScaleWrapperOperator(prec, oscale, inverse_scale=True), noise)
    y = dumbbell["train_y"].double().to(DEV)
    with mgp.settings.max_cholesky_size(2000):
        loss = 0.5 * sum([torch.dot(y, op.matmul(y.view(-1, 1)).squeeze()), -op.inv_quad_logdet(logdet=True)[1],
                          y.shape[0] * math.log(2 * math.pi)])
    grads = torch.autograd.grad(loss, [e, kp, oscale, noise])
    ref = float(g[f"{tag}_nll"])
    assert abs(loss.item() - ref) < 1e-8 * abs(ref)
    for a, b in zip(grads, g[f"{tag}_nll_grads"]):
        assert abs(a.item() - b) < 1e-6 * max(1.0, abs(b))


def test_spmm_autograd_rhs_and_values_vs_oracle():
    import manifold_gp_b200 as mgp
    x = oracle.datasets.torus(4000, seed=3)
    oidx, oval = oracle.knn_graph(x, 10)
    n = x.shape[0]
    for normalization in NORMALIZATIONS:
        eo = torch.tensor(0.3, dtype=torch.float64, requires_grad=True)
        ko = torch.tensor(0.8, dtype=torch.float64, requires_grad=True)
        vo = torch.randn(n, 3, dtype=torch.float64, requires_grad=True)
        w = torch.randn(n, 3, dtype=torch.float64)
        olap = oracle.LaplacianOracle(oval.double(), oidx, n, eo, normalization, True)
        lo = (w * oracle.precision_matmul(olap, 2, ko, vo)).sum()
        go = torch.autograd.grad(lo, [eo, ko, vo])
        e = torch.tensor([[0.3]], dtype=torch.float64, device=DEV, requires_grad=True)
        kp = torch.tensor([[0.8]], dtype=torch.float64, device=DEV, requires_grad=True)
        v = vo.detach().to(DEV).requires_grad_(True)
        lap = mgp.GraphLaplacianOperator(oval.double().to(DEV), oidx.to(DEV), n, e, normalization)
        prec = mgp.PrecisionMaternOperator(lap, 2, kp)
        l = (w.to(DEV) * prec.matmul(v)).sum()
        gg = torch.autograd.grad(l, [e, kp, v])
        assert abs(l.item() - lo.item()) < 1e-9 * abs(lo.item())
        assert abs(gg[0].item() - go[0].item()) < 1e-7 * max(1.0, abs(go[0].item()))
        assert abs(gg[1].item() - go[1].item()) < 1e-7 * max(1.0, abs(go[1].item()))
        assert rel_err(gg[2], go[2]) < 1e-9


@pytest.mark.parametrize("normalization", NORMALIZATIONS)
def test_out_of_sample_vs_reference_dense(golden_k10, normalization):
    """test_outofsample (test/_test_functions.py:134-164)."""
    g = golden_k10
    tag = gtag(0.5, 1.3, 2, normalization, True)
    for dtype, tol in ((torch.float64, 1e-9), (torch.float32, 1e-5)):
        lap, _, _, _ = _ops(g, 0.5, 1.3, 2, normalization, True, dtype)
        U = torch.from_numpy(g[f"{tag}_evecs"]).to(dtype).to(DEV)
        ev = torch.from_numpy(g["oos_edge_value"]).to(dtype).to(DEV)
        ei = torch.from_numpy(g["oos_edge_index"]).long().to(DEV)
        ext = lap.out_of_sample(U, ev, ei)
        assert rel_err(ext, g[f"{tag}_ext_evecs"][:, :6]) < tol


@pytest.mark.parametrize("normalization", NORMALIZATIONS)
def test_schur_vs_reference_dense(golden_k10, normalization):
    import manifold_gp_b200 as mgp
    g = golden_k10
    tag = gtag(0.5, 1.3, 2, normalization, True)
    _, prec, _, _ = _ops(g, 0.5, 1.3, 2, normalization, True, torch.float64)
    mask = torch.from_numpy(g["mask"]).to(DEV)
    V = torch.from_numpy(g["V"]).to(DEV)
    sch = mgp.SchurComplementOperator(prec, mask)
    assert sch.shape == (200, 200)
    with mgp.settings.max_cholesky_size(2000):          # inner solve by dense Cholesky
        assert rel_err(sch.matmul(V[mask]), g[f"{tag}_PschurV"]) < 1e-9
    with mgp.settings.max_cholesky_size(0), mgp.settings.cg_tolerance(1e-6), mgp.settings.max_cg_iterations(2000):
        assert rel_err(sch.matmul(V[mask]), g[f"{tag}_PschurV"]) < 1e-4   # inner solve by CUDA CG


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("normalization", NORMALIZATIONS)
def test_fused_wrapper_paths_vs_reference_dense(golden_k10, normalization, dtype):
    """The solver drivers' buffer interface of the wrappers (``_mgp_matvec`` of Noise(Scale(Precision)), default on since round 2)
    against the reference's dense matern_noisy_precision (golden PnoisyV; noise_wrapper_operator.py:21-22,
    scale_wrapper_operator.py:27-28), its fused dot product, and a CUDA-CG solve through that path against a dense solve."""
    import manifold_gp_b200 as mgp
    g = golden_k10
    tag = gtag(0.5, 1.3, 2, normalization, True)
    tol = TOL[dtype]
    _, prec, _, _ = _ops(g, 0.5, 1.3, 2, normalization, True, dtype)
    V = torch.from_numpy(g["V"]).to(dtype).to(DEV)
    pdiv = mgp.ScaleWrapperOperator(prec, torch.tensor(1.7, dtype=dtype, device=DEV), inverse_scale=True)
    pn = mgp.NoiseWrapperOperator(pdiv, torch.tensor(0.02, dtype=dtype, device=DEV))
    assert pdiv._native() and pn._native()
    st = pn._mgp_structure()
    c = V.shape[1]
    ld = (c + 3) // 4 * 4
    x = torch.zeros(V.shape[0], ld, dtype=dtype, device=DEV)
    x[:, :c] = st.to_internal(V)
    out, tmp = torch.zeros_like(x), torch.zeros_like(x)
    dot = torch.zeros(c, dtype=dtype, device=DEV)
    pn._mgp_matvec(x, out, tmp, dot_with=x, dot_out=dot, ncols=c)
    ref = torch.from_numpy(g[f"{tag}_PnoisyV"]).to(DEV)
    assert rel_err(st.to_external(out[:, :c]), ref) < tol * 10
    assert rel_err(dot, (V.double() * ref.double()).sum(0)) < tol * 100
    pm = mgp.ScaleWrapperOperator(prec, torch.tensor(1.7, dtype=dtype, device=DEV))
    pm._mgp_matvec(x, out, tmp, dot_with=x, dot_out=dot, ncols=c)
    assert rel_err(st.to_external(out[:, :c]), g[f"{tag}_PmulV"]) < tol * (1 if dtype == torch.float64 else 2)
    # CG on the wrapped operator through the fused interface vs a dense solve of the same operator
    dense = pn.to_dense().double()
    want = torch.linalg.solve(dense, V.double())
    with mgp.settings.max_cholesky_size(0), mgp.settings.cg_tolerance(1e-6), mgp.settings.max_cg_iterations(3000):
        got = pn.solve(V)
    assert rel_err(got, want) < (1e-5 if dtype == torch.float64 else 2e-3)


def test_morton_permutation_equivalence():
    """The internal space-filling-curve reordering (attached by NearestNeighbors.graph) must not change any result:
    same operator built with and without the hint, Laplacian + precision + CG + per-node arrays."""
    import manifold_gp_b200 as mgp
    from manifold_gp_b200 import graph, solvers
    x = oracle.datasets.torus(40000, seed=8).to(DEV)
    idx, val = mgp.NearestNeighbors(x).graph(12)
    assert getattr(idx, graph._PERM_ATTR, None) is not None
    perm = getattr(idx, graph._PERM_ATTR)
    assert torch.equal(torch.sort(perm).values, torch.arange(x.shape[0], device=DEV))
    idx_plain = idx.clone()                       # no hint -> identity order
    n = x.shape[0]
    for normalization in NORMALIZATIONS:
        eps = torch.tensor([[0.1]], device=DEV)
        kap = torch.tensor([[0.6]], device=DEV)
        la = mgp.GraphLaplacianOperator(val, idx, n, eps, normalization)
        lb = mgp.GraphLaplacianOperator(val, idx_plain, n, eps, normalization)
        assert la.structure.perm is not None and lb.structure.perm is None
        V = torch.randn(n, 8, device=DEV)
        assert rel_err(la.matmul(V), lb.matmul(V)) < 1e-6
        assert rel_err(la.T.matmul(V), lb.T.matmul(V)) < 1e-6
        assert rel_err(la.degree_mat, lb.degree_mat) < 1e-6
        assert rel_err(la.laplacian_diag, lb.laplacian_diag) < 1e-6
        assert rel_err(la.laplacian_triu, lb.laplacian_triu) < 1e-6
        pa = mgp.PrecisionMaternOperator(la, 2, kap)
        pb = mgp.PrecisionMaternOperator(lb, 2, kap)
        assert rel_err(pa.matmul(V), pb.matmul(V)) < 1e-6
        xa = solvers.linear_cg(pa, V, tolerance=1e-4, max_iter=500)
        xb = solvers.linear_cg(pb, V, tolerance=1e-4, max_iter=500)
        assert rel_err(xa, xb) < 1e-3
