"""Both SpMM kernels (v1 CSR sub-warp, v2 tile-compacted shared-memory) through the C ABI: every column-width
instantiation, fp32/fp64, shift / pre / post, caller-order translation (xmap / ymap), dot epilogue -- against a dense
torch reference built from the same CSR values, and against each other."""
import pytest
import torch

import oracle
from conftest import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def problem():
    import manifold_gp_b200 as mgp
    x = oracle.datasets.torus(20000, seed=12).to(DEV)
    idx, val = mgp.NearestNeighbors(x).graph(16)
    return x, idx, val


def _dense(st, a, diag, n):
    rows = torch.repeat_interleave(torch.arange(n, device=DEV), (st.rowptr[1:] - st.rowptr[:-1]).long())
    A = torch.zeros(n, n, dtype=torch.float64, device=DEV)
    A.index_put_((rows, st.col.long()), a.double(), accumulate=True)
    return torch.diag(diag.double()), A


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_tiled_and_csr_kernels_agree_with_dense(problem, dtype):
    import manifold_gp_b200 as mgp
    from manifold_gp_b200 import graph
    x, idx, val = problem
    n = 6000                                                 # dense reference: keep it small
    xs = x[:n].contiguous()
    idx6, val6 = mgp.NearestNeighbors(xs).graph(12)
    lap = mgp.GraphLaplacianOperator(val6.to(dtype), idx6, n, torch.tensor([[0.15]], dtype=dtype, device=DEV), "symmetric")
    st = lap.structure
    assert st.perm is not None and st.build_tiles() is not None
    _, deg, diag, a = lap._values()
    D, A = _dense(st, a, diag, n)
    tol = 1e-5 if dtype == torch.float32 else 1e-12
    gen = torch.Generator(device=DEV).manual_seed(0)
    pre = torch.rand(n, dtype=dtype, device=DEV, generator=gen) + 0.5
    post = torch.rand(n, dtype=dtype, device=DEV, generator=gen) + 0.5
    shift = torch.tensor([3.7], dtype=dtype, device=DEV)
    for c in (1, 2, 3, 4, 5, 8, 12, 16, 17, 24, 32, 40):
        X = torch.randn(n, c, dtype=dtype, device=DEV, generator=gen)
        for use_pre, use_post, use_shift in ((False, False, False), (True, True, True), (True, False, True)):
            Xd = X.double() * (pre.double().unsqueeze(1) if use_pre else 1.0)
            ref = (D + (float(shift) if use_shift else 0.0) * torch.eye(n, dtype=torch.float64, device=DEV)) @ Xd - A @ Xd
            if use_post:
                ref = ref * post.double().unsqueeze(1)
            outs = {}
            for kern in ("csr", "tiled"):
                graph.SPMM_KERNEL = kern
                try:
                    dot = torch.zeros(c, dtype=dtype, device=DEV)
                    Y = graph.lap_spmm(st, a, diag, X, shift=shift if use_shift else None, pre=pre if use_pre else None,
                                       post=post if use_post else None, dot_with=X, dot_out=dot)
                finally:
                    graph.SPMM_KERNEL = "auto"
                assert rel_err(Y, ref) < tol, (kern, c, use_pre, use_post)
                assert rel_err(dot, (X.double() * ref).sum(0)) < tol * 10, (kern, c)
                outs[kern] = Y
            assert rel_err(outs["tiled"], outs["csr"]) < tol
            vecw = 4 if dtype == torch.float32 else 2
            if not use_pre and c % vecw == 0:
                # v4 pipelined kernel (padded entry streams, vectorised index / value loads)
                graph.SPMM_KERNEL = "pipe"
                try:
                    dot = torch.zeros(c, dtype=dtype, device=DEV)
                    Y = graph.lap_spmm(st, a, diag, X, shift=shift if use_shift else None, post=post if use_post else None,
                                       dot_with=X, dot_out=dot)
                    Y2 = graph.lap_spmm(st, a, diag, X, shift=shift if use_shift else None, post=post if use_post else None)
                finally:
                    graph.SPMM_KERNEL = "auto"
                assert rel_err(Y, ref) < tol, ("pipe", c, use_post)
                assert torch.equal(Y, Y2)
                assert rel_err(dot, (X.double() * ref).sum(0)) < tol * 10, ("pipe", c)
            if not use_pre and c % (16 if dtype == torch.float32 else 8) == 0:
                # v5 warp-interleaved kernel (64-byte rows of X, entry streams in lane-consumption order) and the
                # one-block-per-tile kernel on the same streams
                # ... and (fp32) the paired-row walk of the same kernel (union lists of spatially adjacent row pairs)
                for kern64 in ("wi",) + (("wp",) if dtype == torch.float32 else ()):
                    graph.SPMM_KERNEL = kern64
                    try:
                        dot = torch.zeros(c, dtype=dtype, device=DEV)
                        Y = graph.lap_spmm(st, a, diag, X, shift=shift if use_shift else None, post=post if use_post else None,
                                           dot_with=X, dot_out=dot)
                        Y2 = graph.lap_spmm(st, a, diag, X, shift=shift if use_shift else None, post=post if use_post else None)
                        dot3 = torch.zeros(c, dtype=dtype, device=DEV)
                        Z = torch.randn(n, c, dtype=dtype, device=DEV, generator=gen)
                        graph.lap_spmm(st, a, diag, X, shift=shift if use_shift else None, post=post if use_post else None,
                                       dot_with=Z, dot_out=dot3)
                    finally:
                        graph.SPMM_KERNEL = "auto"
                    assert rel_err(Y, ref) < tol, (kern64, c, use_post)
                    assert graph.LAST_SPMM_KERNEL == ("lap_spmm_wi_kernel<pair>" if kern64 == "wp" else "lap_spmm_wi_kernel")
                    assert torch.equal(Y, Y2)
                    assert rel_err(dot, (X.double() * ref).sum(0)) < tol * 10, (kern64, c)
                    assert (dot3.double() - (Z.double() * ref).sum(0)).abs().max() < tol * 10 * (Z.double().norm() * ref.norm()) / c ** 0.5
        # caller-order translation: x in external order, y in external order
        Xe = X
        ref_ext = st.to_external((D @ st.to_internal(Xe).double()) - A @ st.to_internal(Xe).double())
        for kern in ("csr", "tiled") + (("pipe",) if c % (4 if dtype == torch.float32 else 2) == 0 else ()) + \
                (("wi",) if c % (16 if dtype == torch.float32 else 8) == 0 else ()) + \
                (("wp",) if dtype == torch.float32 and c % 16 == 0 else ()):
            graph.SPMM_KERNEL = kern
            try:
                Y = graph.lap_spmm(st, a, diag, Xe, x_external=True, y_external=True)
                Yi = graph.lap_spmm(st, a, diag, Xe, x_external=True, y_external=False)
            finally:
                graph.SPMM_KERNEL = "auto"
            assert rel_err(Y, ref_ext) < tol
            assert rel_err(st.to_external(Yi), ref_ext) < tol


def test_tiled_kernel_on_unordered_graph_falls_back_or_matches(golden_k10):
    """A graph without the reordering hint: the tile structure is still valid (the dumbbell nodes are ordered along the
    curve); whichever kernel 'auto' picks must match the CSR kernel."""
    import manifold_gp_b200 as mgp
    from manifold_gp_b200 import graph
    g = golden_k10
    idx = torch.from_numpy(g["idx"]).long().to(DEV)
    val = torch.from_numpy(g["val"]).to(DEV)
    n = g["V"].shape[0]
    lap = mgp.GraphLaplacianOperator(val, idx, n, torch.tensor([[0.5]], device=DEV), "symmetric")
    V = torch.from_numpy(g["V"]).float().to(DEV)
    auto = lap.matmul(V)
    graph.SPMM_KERNEL = "csr"
    try:
        ref = lap.matmul(V)
    finally:
        graph.SPMM_KERNEL = "auto"
    assert rel_err(auto, ref) < 1e-6


def test_tile_statistics_cfgc_like(problem):
    """Halo size of a Morton-ordered 2-D manifold: the tile's distinct X rows stay within a few x TILE_ROWS."""
    import manifold_gp_b200 as mgp
    x, idx, val = problem
    lap = mgp.GraphLaplacianOperator(val, idx, x.shape[0], torch.tensor([[0.1]], device=DEV), "symmetric")
    t = lap.structure.build_tiles()
    assert t is not None
    ntiles = (x.shape[0] + t["rows"] - 1) // t["rows"]
    mean_halo = t["halo_total"] / ntiles
    assert mean_halo < 6 * t["rows"], mean_halo
    assert lap.structure.tiled_ok(torch.float32, 16)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_single_column_tile_spmv_matches_dense(problem, dtype):
    """lap_spmv_tile_kernel (one right-hand side: Lanczos, single-RHS CG) against the dense operator, in the structure's
    order, in the caller's order (xmap / ymap) and on a strided column of a wider buffer."""
    import manifold_gp_b200 as mgp
    from manifold_gp_b200 import graph
    x, idx, val = problem
    n = 6000
    xs = x[:n].contiguous()
    idx6, val6 = mgp.NearestNeighbors(xs).graph(12)
    lap = mgp.GraphLaplacianOperator(val6.to(dtype), idx6, n, torch.tensor([[0.15]], dtype=dtype, device=DEV), "symmetric")
    st = lap.structure
    _, deg, diag, a = lap._values()
    D, A = _dense(st, a, diag, n)
    tol = 1e-5 if dtype == torch.float32 else 1e-12
    gen = torch.Generator(device=DEV).manual_seed(3)
    post = torch.rand(n, dtype=dtype, device=DEV, generator=gen) + 0.5
    shift = torch.tensor([2.3], dtype=dtype, device=DEV)
    wide = torch.randn(n, 8, dtype=dtype, device=DEV, generator=gen)
    for X in (torch.randn(n, 1, dtype=dtype, device=DEV, generator=gen), wide[:, 3:4]):
        ref = ((D + float(shift) * torch.eye(n, dtype=torch.float64, device=DEV)) @ X.double() - A @ X.double()) * post.double().unsqueeze(1)
        graph.SPMM_KERNEL = "spmv"
        try:
            Y = graph.lap_spmm(st, a, diag, X, shift=shift, post=post)
            assert graph.LAST_SPMM_KERNEL == "lap_spmv_tile_kernel"
            # caller's row order in and out: P^T M P applied to the un-permuted vector
            Xe = st.to_external(X.contiguous())
            Ye = graph.lap_spmm(st, a, diag, Xe, shift=shift, post=post, x_external=True, y_external=True)
        finally:
            graph.SPMM_KERNEL = "auto"
        assert rel_err(Y, ref) < tol
        assert rel_err(st.to_internal(Ye), ref) < tol
        # fused dot epilogue: dot_with = x itself, and another vector with the same row stride
        graph.SPMM_KERNEL = "spmv"
        try:
            dot = torch.zeros(1, dtype=dtype, device=DEV)
            Yd = graph.lap_spmm(st, a, diag, X, shift=shift, post=post, dot_with=X, dot_out=dot)
            Z = torch.randn(n, 8, dtype=dtype, device=DEV, generator=gen)[:, 2:3] if X.stride(0) == 8 else \
                torch.randn(n, 1, dtype=dtype, device=DEV, generator=gen)
            dot2 = torch.zeros(1, dtype=dtype, device=DEV)
            graph.lap_spmm(st, a, diag, X, shift=shift, post=post, dot_with=Z, dot_out=dot2)
            assert graph.LAST_SPMM_KERNEL == "lap_spmv_tile_kernel"
        finally:
            graph.SPMM_KERNEL = "auto"
        assert torch.equal(Yd, Y)
        assert rel_err(dot, (X.double() * ref).sum(0)) < tol * 10
        assert (dot2.double() - (Z.double() * ref).sum(0)).abs().max() < tol * 10 * Z.double().norm() * ref.norm()
    # the operator-level matvec (auto dispatch) uses it for one column
    v = torch.randn(n, dtype=dtype, device=DEV, generator=gen)
    out = lap._matmul(v)
    assert graph.LAST_SPMM_KERNEL == "lap_spmv_tile_kernel"
    assert rel_err(out, (lap.to_dense().double() @ v.double())) < tol


def test_paired_walk_on_ragged_graphs_and_epilogues(problem):
    """The paired-row walk (fp32) on graphs whose last tile is partial and whose rows have very different lengths (k = 5 and
    k = 40), against the single-row walk of the same kernel and a float64 CSR reference; wrapper epilogue (y = add + coef * Ax),
    the CG done flag, and many more tiles than thread blocks (byte-ring wrap-around)."""
    import manifold_gp_b200 as mgp
    from manifold_gp_b200 import graph, _lib
    x, _, _ = problem
    gen = torch.Generator(device=DEV).manual_seed(3)
    for n, k in ((20000, 40), (19999, 5), (130, 6), (100, 3)):
        xs = x[:n].contiguous()
        idx, val = mgp.NearestNeighbors(xs).graph(k)
        lap = mgp.GraphLaplacianOperator(val, idx, n, torch.tensor([[0.12]], device=DEV), "symmetric")
        st = lap.structure
        _, deg, diag, a = lap._values()
        assert st.pair_tiles() is not None
        for c in (16, 48):
            X = torch.randn(n, c, device=DEV, generator=gen)
            Z = torch.randn(n, c, device=DEV, generator=gen)
            rows = torch.repeat_interleave(torch.arange(n, device=DEV), (st.rowptr[1:] - st.rowptr[:-1]).long())
            ref = diag.double().unsqueeze(1) * X.double()
            ref.index_add_(0, rows, -a.double().unsqueeze(1) * X.double()[st.col.long()])
            res = {}
            for kern in ("wi", "wp"):
                graph.SPMM_KERNEL = kern
                try:
                    dot = torch.zeros(c, device=DEV)
                    Y = graph.lap_spmm(st, a, diag, X, dot_with=Z, dot_out=dot)
                    coef = torch.tensor([0.37], device=DEV)
                    Ye = graph.lap_spmm(st, a, diag, X, ep_coef=coef, ep_add=Z)
                    done = torch.ones(1, device=DEV)
                    keep = torch.full((n, c), 7.0, device=DEV)
                    graph.lap_spmm(st, a, diag, X, out=keep, done_flag=done)
                finally:
                    graph.SPMM_KERNEL = "auto"
                assert rel_err(Y, ref) < 1e-5, (kern, n, k, c)
                assert rel_err(Ye, Z.double() + 0.37 * ref) < 1e-5, (kern, n, k, c)
                assert bool((keep == 7.0).all())
                assert (dot.double() - (Z.double() * ref).sum(0)).abs().max() < 1e-4 * (Z.double().norm() * ref.norm()) / c ** 0.5
                res[kern] = Y
            assert rel_err(res["wp"], res["wi"]) < 2e-6


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_sddmm_all_column_widths_against_torch(problem, dtype):
    """mgp_lap_sddmm (backward of the SpMM w.r.t. the matrix entries; both its scalar and its 128-bit instantiations) against the
    definition evaluated with torch ops: g_a[p] = -post_i pre_j <gy_i, x_j>, g_diag[i] = post_i pre_i <gy_i, x_i>."""
    import manifold_gp_b200 as mgp
    from manifold_gp_b200 import graph
    x, idx, val = problem
    n = x.shape[0]
    lap = mgp.GraphLaplacianOperator(val.to(dtype), idx, n, torch.tensor([[0.15]], dtype=dtype, device=DEV), "symmetric")
    st = lap.structure
    gen = torch.Generator(device=DEV).manual_seed(9)
    rows = torch.repeat_interleave(torch.arange(n, device=DEV), (st.rowptr[1:] - st.rowptr[:-1]).long())
    cols = st.col.long()
    pre = torch.rand(n, dtype=dtype, device=DEV, generator=gen) + 0.5
    post = torch.rand(n, dtype=dtype, device=DEV, generator=gen) + 0.5
    tol = 1e-5 if dtype == torch.float32 else 1e-12
    for c in (1, 2, 3, 4, 8, 11, 12, 16, 24, 32):
        X = torch.randn(n, c, dtype=dtype, device=DEV, generator=gen)
        G = torch.randn(n, c, dtype=dtype, device=DEV, generator=gen)
        for use in (False, True):
            g_a, g_d = graph.lap_sddmm(st, G, X, pre=pre if use else None, post=post if use else None)
            pi = post.double() if use else torch.ones(n, dtype=torch.float64, device=DEV)
            pj = pre.double() if use else torch.ones(n, dtype=torch.float64, device=DEV)
            ref_a = -(pi[rows] * pj[cols]) * (G.double()[rows] * X.double()[cols]).sum(1)
            ref_d = pi * pj * (G.double() * X.double()).sum(1)
            assert rel_err(g_a, ref_a) < tol, (c, use)
            assert rel_err(g_d, ref_d) < tol, (c, use)
