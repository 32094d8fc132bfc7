#!/usr/bin/env python
"""Partitioned graph construction under torchrun (2+ GPUs): every rank searches / symmetrises / builds structure and values for
its rows only (distributed.PartitionedGraph); the partitioned solve is compared with the replicated single-GPU operators that
rank 0 builds for the check.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 profiles/dist_check_partitioned.py [n]"""
import json, os, sys, time, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import manifold_gp_b200 as mgp
from manifold_gp_b200 import distributed as D, solvers, settings
from manifold_gp_b200.utils import synthetic

warnings.simplefilter("ignore")
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
dev = torch.device("cuda", local); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
k, nu, kappa, c, tol = 32, 2, 0.5, 16, 1e-6
x = synthetic.torus(n, seed=0, device=dev)
torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
pg = D.PartitionedGraph(x, k)
torch.cuda.synchronize(); dist.barrier(); t_build = time.perf_counter() - t0
kd = pg.gather(pg.kth_dist2.unsqueeze(1)).squeeze(1)
eps = float(kd.sqrt().median())
t0 = time.perf_counter()
op = D.PartitionedPrecision(pg, eps, nu, kappa)
torch.cuda.synchronize(); dist.barrier(); t_vals = time.perf_counter() - t0
B = torch.randn(n, c, device=dev, generator=torch.Generator(device=dev).manual_seed(1))
b_loc = pg.to_local(B).contiguous()
cg = D.PeerCG(op, c, torch.float32, tolerance=tol, max_iter=4000)
xs, info = cg.solve_polished(b_loc)
torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
xs, info = cg.solve_polished(b_loc)
torch.cuda.synchronize(); dist.barrier(); t_solve = time.perf_counter() - t0
sol = pg.gather(xs)
mem = torch.cuda.max_memory_allocated() / 2 ** 30
out = None
if rank == 0:
    idx, val = mgp.NearestNeighbors(x).graph(k)
    lap = mgp.GraphLaplacianOperator(val, idx, n, torch.tensor([[eps]], device=dev), "symmetric", True)
    prec = mgp.PrecisionMaternOperator(lap, nu, torch.tensor([[kappa]], device=dev))
    ref, rinfo = solvers.linear_cg(prec, B, tolerance=tol, max_iter=4000, return_info=True)
    err = (sol - ref).double().norm(dim=0) / ref.double().norm(dim=0)
    true_rel = float(((prec.matmul(sol) - B).double().norm(dim=0) / B.double().norm(dim=0)).mean())
    out = {"n": n, "k": k, "world": world, "rows_rank0": pg.n_loc, "halo_rows_rank0": pg.n_ext - pg.n_loc, "entries_rank0": pg.st.nnz,
           "entries_global": lap.structure.nnz, "build_s": round(t_build, 3), "values_and_streams_s": round(t_vals, 3),
           "solve_ms": round(t_solve * 1e3, 1), "iterations": info["iterations"], "polish": info.get("polish"),
           "single_gpu_iterations": int(rinfo["iterations"]), "solution_rel_diff_max": float(err.max()),
           "true_relative_residual": true_rel, "peak_mem_GiB_rank0_before_reference_build": round(mem, 2)}
    print(json.dumps(out))
dist.barrier()
dist.destroy_process_group()
