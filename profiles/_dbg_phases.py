import os, sys, json
sys.path.insert(0, "/root/repo")
import torch
import manifold_gp_b200 as mgp
from manifold_gp_b200 import graph
from manifold_gp_b200.utils import synthetic
n=1_000_000; dev=torch.device("cuda:0")
x = synthetic.torus(n, seed=0, device=dev)
idx, val = mgp.NearestNeighbors(x).graph(32)
lap = mgp.GraphLaplacianOperator(val, idx, n, torch.tensor([[0.027417]], device=dev), "symmetric", True)
st = lap.structure; _, _, diag, a = lap._values()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
def timeit(fn, reps=30, warm=5):
    for _ in range(warm): fn()
    ev0.record()
    for _ in range(reps): fn()
    ev1.record(); torch.cuda.synchronize()
    return ev0.elapsed_time(ev1)*1e3/reps
graph.SPMM_KERNEL="tiled"
for c in (1,16):
    P = torch.randn(n, c, device=dev); V = torch.zeros_like(P)
    out={}
    for dbg in ("0","1","3"):
        os.environ["MGP_TILED_DEBUG"]=dbg
        out[dbg]=round(timeit(lambda: graph.lap_spmm(st, a, diag, P, out=V)),1)
    print(c, out)
