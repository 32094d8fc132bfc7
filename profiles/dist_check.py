#!/usr/bin/env python
"""Development check of the multi-GPU CG transports (torchrun, >= 2 GPUs): NCCL path vs peer-memory path on the same
problem -- iteration counts, solutions, residual, time per iteration."""
import os, sys, time, json, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import manifold_gp_b200 as mgp
from manifold_gp_b200 import graph, distributed as D
from manifold_gp_b200.utils import synthetic

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", local); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
warnings.simplefilter("ignore")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
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
res = {}
sols = {}
for name, cls in (("nccl", D.DistCG), ("peer", D.PeerCG)):
    cg = cls(op, c, torch.float32, tolerance=1e-6, max_iter=4000)
    xs, info = cg.solve(b_loc)
    torch.cuda.synchronize(); dist.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(); xs, info = cg.solve(b_loc); ev1.record(); torch.cuda.synchronize(); dist.barrier()
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev); dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    x_all = [torch.empty((part.range(r)[1] - part.range(r)[0], c), device=dev) for r in range(world)]
    dist.all_gather(x_all, xs.contiguous())
    sol = gst.to_external(torch.cat(x_all))
    rel = float(((prec.matmul(sol) - B).double().norm(dim=0) / B.double().norm(dim=0)).mean())
    res[name] = dict(ms=round(float(ms), 2), iters=info["iterations"], us_per_iter=round(float(ms) * 1e3 / max(info["iterations"], 1), 1),
                     converged=info["converged"], true_rel=rel)
    sols[name] = sol
# micro-timings of the building blocks (every rank enqueues the same sequence)
def timeit(fn, reps=200):
    for _ in range(10): fn()
    torch.cuda.synchronize(); dist.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(reps): fn()
    ev1.record(); torch.cuda.synchronize()
    return round(ev0.elapsed_time(ev1) * 1e3 / reps, 2)
res["us_peer_barrier"] = timeit(cg._barrier)
res["us_peer_scalars_pap"] = timeit(lambda: cg._scalars(3, cg.rbuf_pap))
res["us_matvec_2spmm_2barrier"] = timeit(cg._matvec, 50)
gph = torch.cuda.CUDAGraph()
with torch.cuda.graph(gph):
    for _ in range(20): cg._barrier()
res["us_peer_barrier_in_graph"] = round(timeit(gph.replay, 20) / 20, 2)
gph2 = torch.cuda.CUDAGraph()
with torch.cuda.graph(gph2):
    for _ in range(20): cg._matvec()
res["us_matvec_in_graph"] = round(timeit(gph2.replay, 10) / 20, 2)
from manifold_gp_b200 import _lib
from manifold_gp_b200._lib import c_int32, c_int64, ptr, stream
def vec():
    _lib.call("mgp_cg_rupdate_f32", ptr(cg.r), ptr(cg.v), c_int64(cg.ld), c_int64(op.n_loc), c_int32(c), ptr(cg.state), None, c_int32(0), ptr(cg.rbuf), ptr(cg.ws), stream())
    _lib.call("mgp_cg_pxupdate_f32", ptr(cg.x), ptr(cg.p), ptr(cg.r), c_int64(cg.ld), c_int64(op.n_loc), c_int32(c), ptr(cg.state), stream())
cg.state.zero_()
gph3 = torch.cuda.CUDAGraph()
with torch.cuda.graph(gph3):
    for _ in range(20): vec()
res["us_rupdate_pxupdate_in_graph"] = round(timeit(gph3.replay, 10) / 20, 2)
res["rows_per_rank"] = op.n_loc; res["halo_rows"] = int(op.plan.halo_ids.numel())
res["rel_diff_peer_vs_nccl"] = float((sols["peer"] - sols["nccl"]).norm() / sols["nccl"].norm())
if rank == 0:
    print(json.dumps(res))
dist.barrier(); dist.destroy_process_group()
