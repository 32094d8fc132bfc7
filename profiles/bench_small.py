#!/usr/bin/env python
"""Per-rank work of the strong-scaling regime on ONE GPU: a cfg-C graph cut down to the rows one of 8 ranks owns (default
125k), everything L2 resident.  Times (CUDA graphs of 64 repetitions, CUDA events) the building blocks of the single-reduction
peer CG: SpMM alone, the nu-launch matvec, the full iteration.   python profiles/bench_small.py [n]"""
import json, os, sys, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import manifold_gp_b200 as mgp
from manifold_gp_b200 import distributed as D, graph, solvers
from manifold_gp_b200.utils import synthetic

warnings.simplefilter("ignore")
dev = torch.device("cuda:0"); torch.cuda.set_device(dev)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29577")
dist.init_process_group("nccl", rank=0, world_size=1, device_id=dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 125_000
k, c, nu, kappa = 32, 16, 2, 0.5
x = synthetic.torus(8 * n, seed=0, device=dev)[:n].contiguous() if False else synthetic.torus(n, seed=0, device=dev)
knn = mgp.NearestNeighbors(x)
idx, val = knn.graph(k)
d2, _ = knn.search(x[:4096].contiguous(), k)
eps = float(d2[:, k - 1].sqrt().median())
lap = mgp.GraphLaplacianOperator(val, idx, n, torch.tensor([[eps]], device=dev), "symmetric", True)
prec = mgp.PrecisionMaternOperator(lap, nu, torch.tensor([[kappa]], device=dev))
gst = lap.structure
_, _, diag, a = lap._values()
part = D.RowPartition(n, 1, align=gst.TILE_ROWS)
op = D.DistPrecision(gst, diag, a, prec._shift(), nu, part, 0)
B = torch.randn(n, c, device=dev)
res = {"n": n, "tiles": (n + 127) // 128}
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def graph_time(fn, reps=64, outer=5):
    g = solvers._capture(fn, reps)
    for _ in range(2):
        g.replay()
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(outer):
        g.replay()
    ev1.record(); torch.cuda.synchronize()
    return round(ev0.elapsed_time(ev1) * 1e3 / (reps * outer), 2)


P = torch.randn(n, c, device=dev); V = torch.empty_like(P); dot = torch.zeros(c, device=dev)
st = lap.structure
shift = prec._shift()
graph.lap_spmm(st, a, diag, P, shift=shift, out=V)
res["us_spmm_wi"] = graph_time(lambda: graph.lap_spmm(st, a, diag, P, shift=shift, out=V))
res["us_spmm_wi_dot"] = graph_time(lambda: graph.lap_spmm(st, a, diag, P, shift=shift, out=V, dot_with=P, dot_out=dot))
for mode in ("cg1", "fused"):
    cg = D.PeerCG(op, c, torch.float32, tolerance=0.0, max_iter=100000, mode=mode)
    cg.solve(gst.to_internal(B).contiguous()) if False else None
    # a solve that never converges inside the timed region: state initialised by the public path, then iterations only
    cg.tol = 0.0
    cg.max_iter = 48
    cg.use_graph = False
    cg.solve(gst.to_internal(B).contiguous())            # 48 eager iterations: buffers / flags / state in a live configuration
    scal = solvers.S_NARR * c
    cg.state[scal + solvers.K_DONE] = 0.0
    cg.state[scal + 8] = 1e9                                # K_MAXITER
    mv = cg._matvec_cg1 if mode == "cg1" else cg._matvec
    res[f"us_matvec_{mode}"] = graph_time(mv)
    res[f"us_iteration_{mode}"] = graph_time(cg._iteration)
    del cg
print(json.dumps(res))
dist.destroy_process_group()
