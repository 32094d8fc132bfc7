"""ncu target: lap_spmm_wi_kernel at cfg-C with 16 fp32 right-hand sides, single-row walk ("wi") then paired-row walk ("wp"):
    ncu --set full --import-source on --clock-control none -k regex:lap_spmm_wi --launch-skip 4 -c 2 python profiles/prof_spmm_walks.py"""
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
prec = mgp.PrecisionMaternOperator(lap, 2, torch.tensor([[0.5]], device=dev))
st = lap.structure
_, _, diag, a = lap._values()
shift = prec._shift()
P = torch.randn(n, 16, device=dev); V = torch.empty_like(P)
order = sys.argv[1].split(",") if len(sys.argv) > 1 else ["wi", "wi", "wp", "wp", "wi", "wp"]
for kern in order:
    graph.SPMM_KERNEL = kern
    graph.lap_spmm(st, a, diag, P, shift=shift, out=V)
graph.SPMM_KERNEL = "auto"
torch.cuda.synchronize()
