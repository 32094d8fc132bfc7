"""ncu target: a few launches of the single-column tile SpMV at cfg-C."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import manifold_gp_b200 as mgp
from manifold_gp_b200 import graph
from manifold_gp_b200.utils import synthetic
n = 1_000_000
dev = torch.device("cuda:0")
x = synthetic.torus(n, seed=0, device=dev)
idx, val = mgp.NearestNeighbors(x).graph(32)
lap = mgp.GraphLaplacianOperator(val, idx, n, torch.tensor([[0.027417]], device=dev), "symmetric", True)
st = lap.structure
_, _, diag, a = lap._values()
p1 = torch.randn(n, 1, device=dev); v1 = torch.empty_like(p1)
for _ in range(6):
    graph.lap_spmm(st, a, diag, p1, out=v1); graph.lap_spmm(st, a, diag, v1, out=p1)
torch.cuda.synchronize()
print(graph.LAST_SPMM_KERNEL)
