"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: row partition, halo plan / exchange, and the placement of
the all-reduces in the partitioned CG -- with the local arithmetic done by the oracle-style torch ops (the CUDA kernels
need a GPU; the same plumbing runs over NCCL in bench.py --gpus N)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from manifold_gp_b200.distributed import HaloPlan, RowPartition
        torch.manual_seed(0)
        n, k = 1500, 8
        x = oracle.datasets.circle_curve(n, seed=3)
        # order the points along the curve (what the Morton order does on the GPU) so that halos are thin
        ang = torch.atan2(x[:, 1], x[:, 0])
        x = x[torch.argsort(ang)].contiguous()
        idx, val = oracle.knn_graph(x, k)
        lap = oracle.LaplacianOracle(val.double(), idx, n, 0.2, "symmetric", True)
        # global CSR over both directions
        rows = torch.cat([idx[0], idx[1]]); cols = torch.cat([idx[1], idx[0]]); a = torch.cat([lap.laplacian_triu] * 2)
        order = torch.argsort(rows * n + cols)
        rows, cols, a = rows[order], cols[order], a[order]
        rowptr = torch.zeros(n + 1, dtype=torch.int64); rowptr[1:] = torch.cumsum(torch.bincount(rows, minlength=n), 0)
        part = RowPartition(n, world, align=128)
        lo, hi = part.range(rank)
        assert part.bounds[0] == 0 and part.bounds[-1] == n and all(b % 128 == 0 for b in part.bounds[:-1])
        p0, p1 = int(rowptr[lo]), int(rowptr[hi])
        plan = HaloPlan(part, rank, cols[p0:p1])
        n_loc, H = hi - lo, plan.halo_ids.numel()
        assert H < n_loc                                   # thin halo for an ordered curve
        lcol = plan.to_local(cols[p0:p1])
        assert int(lcol.max()) < n_loc + H
        # ---- halo exchange: the extended local vector equals the global vector at [own | halo] rows -----------------
        g = torch.Generator().manual_seed(1)
        X = torch.randn(n, 3, generator=g, dtype=torch.float64)
        xe = torch.zeros(n_loc + H, 3, dtype=torch.float64)
        xe[:n_loc] = X[lo:hi]
        plan.exchange(xe)
        assert torch.equal(xe[n_loc:], X[plan.halo_ids])
        # ---- partitioned matvec == global matvec ------------------------------------------------------------------------
        lrows = rows[p0:p1] - lo
        y_loc = lap.laplacian_diag[lo:hi].unsqueeze(1) * xe[:n_loc]
        y_loc.index_add_(0, lrows, -(a[p0:p1].unsqueeze(1) * xe[lcol]))
        assert torch.allclose(y_loc, lap.matmul(X)[lo:hi], rtol=1e-12, atol=1e-12)
        # ---- partitioned CG with all-reduced dots == oracle CG ------------------------------------------------------------
        shift = 2 * 2 / 0.7 ** 2

        def local_prec(p_ext):                               # (shift + L)^2 on my rows; p_ext has halo rows filled in
            t = torch.zeros_like(p_ext)
            t[:n_loc] = (lap.laplacian_diag[lo:hi] + shift).unsqueeze(1) * p_ext[:n_loc]
            t[:n_loc].index_add_(0, lrows, -(a[p0:p1].unsqueeze(1) * p_ext[lcol]))
            plan.exchange(t)
            v = (lap.laplacian_diag[lo:hi] + shift).unsqueeze(1) * t[:n_loc]
            v.index_add_(0, lrows, -(a[p0:p1].unsqueeze(1) * t[lcol]))
            return v

        B = X / X.norm(dim=0, keepdim=True)
        xs = torch.zeros(n_loc, 3, dtype=torch.float64); r = B[lo:hi].clone()
        p = torch.zeros(n_loc + H, 3, dtype=torch.float64); p[:n_loc] = r
        rz = (r * r).sum(0); dist.all_reduce(rz)
        for it in range(60):
            plan.exchange(p)
            v = local_prec(p)
            pap = (p[:n_loc] * v).sum(0); dist.all_reduce(pap)
            alpha = rz / pap
            xs += alpha * p[:n_loc]; r -= alpha * v
            rz_new = (r * r).sum(0); dist.all_reduce(rz_new)
            p[:n_loc] = r + (rz_new / rz) * p[:n_loc]
            rz = rz_new
        ref = torch.linalg.solve(oracle.dense_from_matmul(lambda t: oracle.precision_matmul(lap, 2, 0.7, t), n), B)
        err = float((xs - ref[lo:hi]).norm() / ref[lo:hi].norm())
        assert err < 1e-8, err
        # ---- single-reduction variant (Chronopoulos-Gear): ONE all-reduce of (r.r, r.Ar) per iteration ---------------------
        # The recurrences planned for the next multi-GPU round: the matvec is applied to r, the
        # product A p is carried by its own recurrence q = A r + beta q, and both inner products travel in one message.
        xs = torch.zeros(n_loc, 3, dtype=torch.float64); r = B[lo:hi].clone()
        pv = torch.zeros(n_loc, 3, dtype=torch.float64); q = torch.zeros(n_loc, 3, dtype=torch.float64)
        rext = torch.zeros(n_loc + H, 3, dtype=torch.float64)
        gamma_old = alpha_old = None
        n_allreduce = 0
        for it in range(60):
            rext[:n_loc] = r
            plan.exchange(rext)
            s_ = local_prec(rext)                                # A r  (one halo exchange per chained SpMM inside)
            red = torch.stack([(r * r).sum(0), (r * s_).sum(0)]); dist.all_reduce(red); n_allreduce += 1
            gamma, delta = red[0], red[1]
            if it == 0:
                beta = torch.zeros_like(gamma); alpha = gamma / delta
            else:
                beta = gamma / gamma_old
                alpha = gamma / (delta - beta * gamma / alpha_old)
            pv = r + beta * pv
            q = s_ + beta * q
            xs += alpha * pv
            r -= alpha * q
            gamma_old, alpha_old = gamma, alpha
        assert n_allreduce == 60
        err = float((xs - ref[lo:hi]).norm() / ref[lo:hi].norm())
        assert err < 1e-8, err
        results[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_partition_halo_exchange_and_partitioned_cg_gloo():
    world = 2
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), results), nprocs=world, join=True)
    assert dict(results) == {0: "ok", 1: "ok"}


def test_row_partition_alignment():
    from manifold_gp_b200.distributed import RowPartition
    for n, w in ((1_000_000, 8), (1000, 4), (129, 2), (70000, 3)):
        p = RowPartition(n, w, align=128)
        assert p.bounds[0] == 0 and p.bounds[-1] == n
        assert all(b % 128 == 0 or b == n for b in p.bounds)
        assert all(p.bounds[i] <= p.bounds[i + 1] for i in range(w))
        ids = torch.tensor([0, n - 1, p.bounds[1] - 1 if p.bounds[1] > 0 else 0, min(p.bounds[1], n - 1)])
        own = p.owner(ids)
        for i, o in zip(ids.tolist(), own.tolist()):
            assert p.bounds[o] <= i < p.bounds[o + 1]


def _worker_symm(rank, world, port, results):
    """partitioned_symmetrize over gloo: every rank's rows of the symmetrised, mean-coalesced graph equal the oracle's
    restatement of nearest_neighbors.py:39-55 on the whole point cloud -- both directions of every edge, diagonal entries
    twice -- and the partitioned three-pass value build with its two halo gathers (restated with torch ops here; the CUDA
    passes are mgp_lap_values_pass) equals the oracle's values on those rows."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from manifold_gp_b200.distributed import HaloPlan, RowPartition, partitioned_symmetrize
        n, k = 900, 7
        g = torch.Generator().manual_seed(5)
        x = torch.randn(n, 3, generator=g)
        x[17] = x[3]; x[400] = x[3]; x[401] = x[3]                  # duplicate points: ties and self matches off column 0
        d2, nbr = oracle.knn_search(x, x, k)
        ref_idx, ref_val = oracle.symmetrize_coalesce(d2, nbr, n, drop_first=True)
        # the reference's rows (both directions; a diagonal COO entry twice)
        rr = torch.cat([ref_idx[0], ref_idx[1]]); rc = torch.cat([ref_idx[1], ref_idx[0]]); rv = torch.cat([ref_val, ref_val])
        o = torch.argsort(rr * n + rc, stable=True); rr, rc, rv = rr[o], rc[o], rv[o]
        part = RowPartition(n, world, align=128)
        lo, hi = part.range(rank)
        rows = torch.arange(lo, hi).repeat_interleave(k - 1)
        cols = nbr[lo:hi, 1:].reshape(-1).to(torch.int64)
        vals = d2[lo:hi, 1:].reshape(-1)
        row, col, val = partitioned_symmetrize(rows, cols, vals, part, rank)
        sel = (rr >= lo) & (rr < hi)
        assert torch.equal(row, rr[sel]) and torch.equal(col, rc[sel])
        assert torch.allclose(val, rv[sel], rtol=1e-6, atol=0)
        assert int((row == col).sum()) == int((rr[sel] == rc[sel]).sum())
        # ---- value build on my rows with two halo gathers ---------------------------------------------------------------
        plan = HaloPlan(part, rank, col)
        n_loc, H = hi - lo, plan.halo_ids.numel()
        lc = plan.to_local(col); lr = row - lo
        eps = 0.6
        W = torch.exp(val.double() / (-4 * eps * eps))
        dt = torch.zeros(n_loc + H, 1, dtype=torch.float64)
        dt[:n_loc, 0] = 1.0
        dt[:n_loc, 0].index_add_(0, lr, W)
        plan.exchange(dt)                                            # gather 1: unnormalised degrees of the halo rows
        at = W / (dt[lr, 0] * dt[lc, 0])
        dg = torch.zeros(n_loc + H, 1, dtype=torch.float64)
        dg[:n_loc, 0] = 1.0 / dt[:n_loc, 0] ** 2
        dg[:n_loc, 0].index_add_(0, lr, at)
        plan.exchange(dg)                                            # gather 2: normalised degrees of the halo rows
        a = at / (dg[lr, 0].sqrt() * dg[lc, 0].sqrt()) / eps ** 2
        lap = oracle.LaplacianOracle(ref_val.double(), ref_idx, n, eps, "symmetric", True)
        assert torch.allclose(dt[:n_loc, 0], lap.degree_unnorm_mat[lo:hi], rtol=1e-12)
        assert torch.allclose(dg[:n_loc, 0], lap.degree_mat[lo:hi], rtol=1e-12)
        # my rows of L X against the oracle (diagonal entries twice, as in the reference's two scatter passes)
        X = torch.randn(n, 2, generator=torch.Generator().manual_seed(6), dtype=torch.float64)
        xe = torch.zeros(n_loc + H, 2, dtype=torch.float64); xe[:n_loc] = X[lo:hi]
        plan.exchange(xe)
        y = lap.laplacian_diag[lo:hi].unsqueeze(1) * xe[:n_loc]
        y.index_add_(0, lr, -(a.unsqueeze(1) * xe[lc]))
        assert torch.allclose(y, lap.matmul(X)[lo:hi], rtol=1e-10, atol=1e-10)
        # ---- rows sorted by length inside tiles across the partition (partitioned_sort_rows) ----------------------------
        from manifold_gp_b200.distributed import partitioned_sort_rows
        perm0 = torch.arange(n)                                       # the rows' original indices before the refinement
        row2, col2, val2, perm2 = partitioned_sort_rows(row, col, val, part, rank, perm0[lo:hi], 128)
        assert torch.equal(torch.sort(perm2).values, torch.arange(n))                  # a permutation of all rows ...
        assert torch.equal(torch.sort(perm2[lo:hi]).values, torch.arange(lo, hi))      # ... that keeps every rank's block
        deg2 = torch.bincount(row2 - lo, minlength=n_loc)
        for t0 in range(0, n_loc, 128):                                                # descending lengths inside every tile
            d = deg2[t0:t0 + 128]
            assert bool((d[1:] <= d[:-1]).all())
        # the renumbered entries are the old ones under perm2: new id g names the row that had id perm2[g] before
        key_old = torch.sort(row * n + col).values
        key_new = torch.sort(perm2[row2] * n + perm2[col2]).values
        assert torch.equal(key_old, key_new)
        o_old = torch.argsort(row * n + col, stable=True); o_new = torch.argsort(perm2[row2] * n + perm2[col2], stable=True)
        assert torch.equal(val[o_old], val2[o_new])
        assert bool(((row2[1:] * n + col2[1:]) >= (row2[:-1] * n + col2[:-1])).all())  # sorted by (row, col)
        results[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_partitioned_symmetrize_and_value_build_gloo():
    world = 2
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker_symm, args=(world, _free_port(), results), nprocs=world, join=True)
    assert dict(results) == {0: "ok", 1: "ok"}
