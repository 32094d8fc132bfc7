"""GPU parity of the multi-GPU CG transports on ONE device (process group of size 1: every flag / reduction / halo path of the
peer-memory kernels runs, with the rank's own memory as the "peer"): the single-reduction (Chronopoulos-Gear) iteration and the
round-1 fused two-reduction iteration against solvers.linear_cg and the oracle's mBCG -- iteration counts, solutions, the
Lanczos tridiagonals recorded from the CG coefficients.  The 2-rank protocol is covered on CPU (tests/test_distributed_cpu.py,
gloo) and on 2 / 8 GPUs by profiles/dist_check2.py (numbers under profiles/)."""
import os
import socket
import warnings

import pytest
import torch

import oracle
from conftest import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def pg1():
    import torch.distributed as dist
    created = False
    if not dist.is_initialized():
        s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
        dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=0, world_size=1,
                                device_id=torch.device("cuda", 0))
        created = True
    yield dist
    if created:
        dist.destroy_process_group()


def _problem(n, k, dtype, nu=2, kappa=0.7, seed=0):
    import manifold_gp_b200 as mgp
    x = oracle.datasets.torus(n, seed=seed)
    knn = mgp.NearestNeighbors(x.to(DEV))
    idx, val = knn.graph(k)
    d2, _ = knn.search(x[:2048].to(DEV).contiguous(), k)
    eps = float(d2[:, k - 1].sqrt().median())
    lap = mgp.GraphLaplacianOperator(val.to(dtype), idx, n, torch.tensor([[eps]], dtype=dtype, device=DEV), "symmetric", True)
    prec = mgp.PrecisionMaternOperator(lap, nu, torch.tensor([[kappa]], dtype=dtype, device=DEV))
    return lap, prec


@pytest.mark.parametrize("dtype,mode", [(torch.float64, "cg1"), (torch.float32, "cg1"), (torch.float64, "fused")])
def test_peer_cg_matches_linear_cg(pg1, dtype, mode):
    from manifold_gp_b200 import distributed as D, solvers
    n, k, c, nu = 20000, 12, 16, 2
    lap, prec = _problem(n, k, dtype, nu=nu)
    gst = lap.structure
    assert gst.tiles is not None and "wptr" in gst.tiles        # the peer path needs the warp-interleaved streams
    _, _, diag, a = lap._values()
    part = D.RowPartition(n, 1, align=gst.TILE_ROWS)
    op = D.DistPrecision(gst, diag, a, prec._shift().to(dtype), nu, part, 0)
    B = torch.randn(n, c, dtype=dtype, device=DEV, generator=torch.Generator(device=DEV).manual_seed(1))
    tol = 1e-6 if dtype == torch.float64 else 1e-5     # (the published eps = 1e-10 masks stall any CG below ~1e-6: |r|^2 < eps)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref, rinfo = solvers.linear_cg(prec, B, tolerance=tol, max_iter=3000, return_info=True)
    cg = D.PeerCG(op, c, dtype, tolerance=tol, max_iter=3000, mode=mode)
    assert cg.mode == mode
    for rep in range(2):                                         # second solve replays the captured graph from the start
        xs, info = cg.solve(gst.to_internal(B).contiguous())
        sol = gst.to_external(xs)
        assert info["converged"]
        if dtype == torch.float64:
            assert info["iterations"] == rinfo["iterations"], (info, dict(rinfo))
            assert rel_err(sol, ref) < 1e-6
        else:
            assert abs(info["iterations"] - rinfo["iterations"]) <= max(3, rinfo["iterations"] // 50)
            assert rel_err(sol, ref) < 2e-3
        true_rel = ((prec._matmul(sol) - B).norm(dim=0) / B.norm(dim=0)).max()
        assert float(true_rel) < (1e-5 if dtype == torch.float64 else 5e-3)


def test_peer_cg1_tridiagonals_match_linear_cg(pg1):
    """The CG coefficients recorded by the single-reduction kernel give the same Lanczos tridiagonals (SLQ log-det input) as
    solvers.linear_cg, which is pinned against the oracle's mBCG in test_gpu_solvers.py."""
    from manifold_gp_b200 import distributed as D, solvers
    dtype = torch.float64
    n, k, c, nu, nt = 12000, 10, 8, 1, 6
    lap, prec = _problem(n, k, dtype, nu=nu, kappa=1.1, seed=3)
    gst = lap.structure
    _, _, diag, a = lap._values()
    part = D.RowPartition(n, 1, align=gst.TILE_ROWS)
    op = D.DistPrecision(gst, diag, a, prec._shift().to(dtype), nu, part, 0)
    B = torch.randn(n, c, dtype=dtype, device=DEV, generator=torch.Generator(device=DEV).manual_seed(5))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref, t_ref, rinfo = solvers.linear_cg(prec, B, n_tridiag=nt, tolerance=1e-4, max_iter=1500, max_tridiag_iter=30, return_info=True)
    cg = D.PeerCG(op, c, dtype, tolerance=1e-4, max_iter=1500, mode="cg1", n_tridiag=nt, max_tridiag_iter=30)
    xs, info = cg.solve(gst.to_internal(B).contiguous())
    assert info["iterations"] == rinfo["iterations"] and info["converged"] == rinfo["converged"]
    t = cg.tridiagonals(info)
    assert t.shape == t_ref.shape
    assert rel_err(t, t_ref) < 1e-8
    assert rel_err(gst.to_external(xs), ref) < 1e-8


def test_dist_backend_routes_training_loss_through_the_peer_cg(pg1):
    """solvers.set_distributed_backend(DistBackend()): inv_quad_logdet of Noise(Scale(Precision)) -- the operator of the training
    loss (riemann_gp.py:32-39, utils/train_model.py:67-68) -- through the row-partitioned single-reduction CG (3 nu SpMM launches
    per matvec with the wrapper algebra in their epilogues) against the single-GPU drivers: same iteration count, inverse
    quadratic form, SLQ log-det (same probes) and the same gradients w.r.t. bandwidth, lengthscale, output scale and noise."""
    import manifold_gp_b200 as mgp
    from manifold_gp_b200 import distributed as D, solvers
    dtype = torch.float64
    n, k, nu = 20000, 12, 1
    x = oracle.datasets.torus(n, seed=4)
    knn = mgp.NearestNeighbors(x.to(DEV))
    idx, val = knn.graph(k)
    y = torch.sin(2.0 * x[:, 0]).to(dtype).to(DEV)
    probes = torch.randn(n, 6, dtype=dtype, device=DEV, generator=torch.Generator(device=DEV).manual_seed(9))

    def run(backend):
        solvers.set_distributed_backend(backend)
        try:
            eps = torch.tensor([[0.2]], dtype=dtype, device=DEV, requires_grad=True)
            kap = torch.tensor([[0.9]], dtype=dtype, device=DEV, requires_grad=True)
            osc = torch.tensor(1.7, dtype=dtype, device=DEV, requires_grad=True)
            noi = torch.tensor(2e-3, dtype=dtype, device=DEV, requires_grad=True)
            lap = mgp.GraphLaplacianOperator(val.to(dtype), idx, n, eps, "symmetric", True)
            prec = mgp.PrecisionMaternOperator(lap, nu, kap)
            op = mgp.NoiseWrapperOperator(mgp.ScaleWrapperOperator(prec, osc, inverse_scale=True), noi)
            with mgp.settings.max_cholesky_size(0), mgp.settings.cg_tolerance(1e-4), mgp.settings.max_cg_iterations(2000):
                iq, ld, info = solvers.inv_quad_logdet(op, inv_quad_rhs=y.unsqueeze(-1), logdet=True, probes=probes, return_info=True)
            (iq + ld).backward()
            return float(iq), float(ld), info["iterations"], [float(t.grad.reshape(-1)[0]) for t in (eps, kap, osc, noi)]
        finally:
            solvers.set_distributed_backend(None)

    want = run(None)
    be = D.DistBackend(min_rows=1000)
    got = run(be)
    assert be.solves >= 1
    assert got[2] == want[2]
    assert abs(got[0] - want[0]) <= 1e-7 * abs(want[0]) and abs(got[1] - want[1]) <= 1e-7 * abs(want[1])
    for a, b in zip(got[3], want[3]):
        assert abs(a - b) <= 1e-5 * max(abs(b), 1e-12), (got, want)


def test_distributed_lanczos_matches_the_single_gpu_driver(pg1):
    """distributed.dist_lanczos_tridiag (row-partitioned vectors, all-reduced re-orthogonalisation dots) on a process group of size 1
    against solvers.lanczos_tridiag with the same start vector: same tridiagonal, same Ritz values / vectors."""
    import manifold_gp_b200 as mgp
    from manifold_gp_b200 import distributed as D, solvers
    dtype = torch.float64
    n, k = 12000, 10
    x = oracle.datasets.torus(n, seed=5)
    idx, val = mgp.NearestNeighbors(x.to(DEV)).graph(k)
    lap = mgp.GraphLaplacianOperator(val.to(dtype), idx, n, torch.tensor([[0.2]], dtype=dtype, device=DEV), "symmetric", True)
    v0 = torch.randn(n, dtype=dtype, device=DEV, generator=torch.Generator(device=DEV).manual_seed(11))
    q_ref, t_ref = solvers.lanczos_tridiag(lap, 40, init_vec=v0)
    be = D.DistBackend(min_rows=1000)
    evals, evecs, t = be.lanczos_eigenpairs(lap, 40, init_vec=v0)
    assert t.shape == t_ref.shape
    assert rel_err(t, t_ref) < 1e-8
    ev_ref, v_ref = solvers.lanczos_tridiag_to_diag(t_ref)
    assert rel_err(evals, ev_ref) < 1e-8
    # Ritz vectors: the same q^T V as the single-GPU driver's (eigenvector signs of eigh(T) may differ between the two T's)
    ref_vecs = q_ref.T @ v_ref
    assert evecs.shape == ref_vecs.shape
    assert rel_err(evecs.abs(), ref_vecs.abs()) < 1e-6


def test_partitioned_solve_with_the_true_residual_correction(pg1):
    """DistCG.solve_polished (the partitioned form of settings.cg_polish): on an ill-conditioned fp32 system the recurrence
    converges while the true residual drifts; one correction solve on r = b - P x brings the TRUE residual within 10 x tol, the
    iteration counts of the main run are unchanged, and the result agrees with the single-GPU polished solve."""
    from manifold_gp_b200 import distributed as D, settings, solvers
    n, k, c, nu = 60000, 16, 16, 2
    lap, prec = _problem(n, k, torch.float32, nu=nu, kappa=0.5)
    gst = lap.structure
    _, _, diag, a = lap._values()
    part = D.RowPartition(n, 1, align=gst.TILE_ROWS)
    op = D.DistPrecision(gst, diag, a, prec._shift(), nu, part, 0)
    B = torch.randn(n, c, device=DEV, generator=torch.Generator(device=DEV).manual_seed(2))
    tol = 1e-6
    cg = D.PeerCG(op, c, torch.float32, tolerance=tol, max_iter=4000)
    b_loc = gst.to_internal(B).contiguous()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        x0, info0 = cg.solve(b_loc)
        x1, info1 = cg.solve_polished(b_loc)
        with settings.cg_polish(True):
            ref, rinfo = solvers.linear_cg(prec, B, tolerance=tol, max_iter=4000, return_info=True)
    assert info0["converged"] and info1["converged"] and info1["iterations"] == info0["iterations"]
    true = lambda xs: float(((prec._matmul(gst.to_external(xs)) - B).double().norm(dim=0) / B.double().norm(dim=0)).mean())
    t0, t1 = true(x0), true(x1)
    assert abs(info1["polish"]["true_residual_before"] - t0) < 0.05 * t0
    if t0 > 8 * tol:                                   # the drift is there: the correction must remove it
        assert info1["polish"]["iterations"] > 0 and t1 <= 10 * tol and t1 < 0.5 * t0
    else:
        assert info1["polish"]["iterations"] == 0
    assert rel_err(gst.to_external(x1), ref) < 1e-4
    # apply(): the partitioned operator outside the iteration equals the single-GPU matvec
    assert rel_err(gst.to_external(cg.apply(b_loc)), prec._matmul(B)) < 1e-5


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_partitioned_graph_construction_matches_the_replicated_operators(pg1, dtype):
    """distributed.PartitionedGraph / PartitionedPrecision (a rank searches, symmetrises, builds structure and values for ITS rows
    only; here the group has one rank, so "its rows" are all rows and every exchange is the degenerate case -- the 2-rank
    protocol is covered by tests/test_distributed_cpu.py over gloo and by profiles/dist_check_partitioned.py on 2 GPUs):
    degrees, diagonal, the matvec and the CG solve against the replicated operators built by NearestNeighbors.graph."""
    import manifold_gp_b200 as mgp
    from manifold_gp_b200 import distributed as D, solvers
    n, k, nu, kappa, c = 30000, 12, 2, 0.7, 16
    x = oracle.datasets.torus(n, seed=4).to(DEV)
    x[100] = x[7]; x[101] = x[7]                                   # duplicate points (self match off column 0)
    pg = D.PartitionedGraph(x, k)
    idx, val = mgp.NearestNeighbors(x).graph(k)
    eps = float(pg.kth_dist2.sqrt().median())
    lap = mgp.GraphLaplacianOperator(val.to(dtype), idx, n, torch.tensor([[eps]], dtype=dtype, device=DEV), "symmetric", True)
    prec = mgp.PrecisionMaternOperator(lap, nu, torch.tensor([[kappa]], dtype=dtype, device=DEV))
    assert pg.st.nnz == lap.structure.nnz and pg.n_loc == n and pg.n_ext == n
    tol = 1e-10 if dtype == torch.float64 else 1e-5
    dt, dg, diag, a = pg.values(eps, True, dtype)
    assert rel_err(pg.gather(dt[:n].unsqueeze(1)).squeeze(1), lap.degree_unnorm_mat) < tol
    assert rel_err(pg.gather(dg[:n].unsqueeze(1)).squeeze(1), lap.degree_mat) < tol
    assert rel_err(pg.gather(diag.unsqueeze(1)).squeeze(1), lap.laplacian_diag) < tol
    op = D.PartitionedPrecision(pg, eps, nu, kappa, True, dtype)
    B = torch.randn(n, c, dtype=dtype, device=DEV, generator=torch.Generator(device=DEV).manual_seed(3))
    cg_tol = 1e-6 if dtype == torch.float64 else 1e-5
    cg = D.PeerCG(op, c, dtype, tolerance=cg_tol, max_iter=3000)
    b_loc = pg.to_local(B).contiguous()
    assert rel_err(pg.gather(cg.apply(b_loc)), prec._matmul(B)) < tol
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref, rinfo = solvers.linear_cg(prec, B, tolerance=cg_tol, max_iter=3000, return_info=True)
        xs, info = cg.solve(b_loc)
    assert info["converged"]
    assert abs(info["iterations"] - rinfo["iterations"]) <= max(3, rinfo["iterations"] // 50)
    assert rel_err(pg.gather(xs), ref) < (1e-6 if dtype == torch.float64 else 2e-3)
    # a new bandwidth re-runs the partitioned value build into the same buffers
    ptr_a, ptr_aw = op.a.data_ptr(), op._aw.data_ptr()
    op.update(1.3 * eps, kappa)
    lap2 = mgp.GraphLaplacianOperator(val.to(dtype), idx, n, torch.tensor([[1.3 * eps]], dtype=dtype, device=DEV), "symmetric", True)
    prec2 = mgp.PrecisionMaternOperator(lap2, nu, torch.tensor([[kappa]], dtype=dtype, device=DEV))
    assert op.a.data_ptr() == ptr_a and op._aw.data_ptr() == ptr_aw
    assert rel_err(pg.gather(cg.apply(b_loc)), prec2._matmul(B)) < tol
