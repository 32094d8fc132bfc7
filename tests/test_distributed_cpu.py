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
