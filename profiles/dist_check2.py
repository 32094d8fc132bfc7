#!/usr/bin/env python
"""Round-2 check of the multi-GPU CG transports (torchrun, >= 1 GPU): NCCL path, round-1 fused peer path and the single-reduction
peer path ("cg1") on the same problem -- iteration counts, solutions vs the 1-GPU solvers.linear_cg, time per iteration.
    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 profiles/dist_check2.py [n]"""
import os, sys, json, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import manifold_gp_b200 as mgp
from manifold_gp_b200 import distributed as D, solvers
from manifold_gp_b200.utils import synthetic

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
dev = torch.device("cuda", local); torch.cuda.set_device(dev)
if "MASTER_ADDR" not in os.environ:
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = "29533"
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
warnings.simplefilter("ignore")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
modes = (sys.argv[2] if len(sys.argv) > 2 else "nccl,fused,cg1").split(",")
k, c, nu, kappa = 32, 16, 2, 0.5
x = synthetic.torus(n, seed=0, device=dev)
knn = mgp.NearestNeighbors(x)
d2, _ = knn.search(x[:4096].contiguous(), k)
eps = float(d2[:, k - 1].sqrt().median())
idx, val = knn.graph(k)
lap = mgp.GraphLaplacianOperator(val, idx, n, torch.tensor([[eps]], device=dev), "symmetric", True)
prec = mgp.PrecisionMaternOperator(lap, nu, torch.tensor([[kappa]], device=dev))
gst = lap.structure
_, _, diag, a = lap._values()
part = D.RowPartition(n, world, align=gst.TILE_ROWS)
op = D.DistPrecision(gst, diag, a, prec._shift(), nu, part, rank)
lo, hi = part.range(rank)
B = torch.randn(n, c, device=dev, generator=torch.Generator(device=dev).manual_seed(1))
b_loc = gst.to_internal(B)[lo:hi].contiguous()
res = {"n": n, "world": world, "rows_per_rank": op.n_loc, "halo_rows": int(op.plan.halo_ids.numel())}
ref, rinfo = solvers.linear_cg(prec, B, tolerance=1e-6, max_iter=4000, return_info=True)     # every rank: the 1-GPU solve
res["single_gpu_iters"] = rinfo["iterations"]
for name in modes:
    cg = D.DistCG(op, c, torch.float32, tolerance=1e-6, max_iter=4000) if name == "nccl" else \
        D.PeerCG(op, c, torch.float32, tolerance=1e-6, max_iter=4000, mode=name)
    xs, info = cg.solve(b_loc)
    torch.cuda.synchronize(); dist.barrier()
    best = None
    for _ in range(3):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(); xs, info = cg.solve(b_loc); ev1.record(); torch.cuda.synchronize(); dist.barrier()
        ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev); dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        best = float(ms) if best is None else min(best, float(ms))
    x_all = [torch.empty((part.range(r)[1] - part.range(r)[0], c), device=dev) for r in range(world)]
    dist.all_gather(x_all, xs.contiguous())
    sol = gst.to_external(torch.cat(x_all))
    rel = float(((prec.matmul(sol) - B).double().norm(dim=0) / B.double().norm(dim=0)).mean())
    res[name] = dict(ms=round(best, 2), iters=info["iterations"], us_per_iter=round(best * 1e3 / max(info["iterations"], 1), 1),
                     converged=info["converged"], true_rel=rel,
                     rel_diff_vs_single_gpu=float((sol - ref).norm() / ref.norm()))
    del cg
if rank == 0:
    print(json.dumps(res))
dist.barrier(); dist.destroy_process_group()
