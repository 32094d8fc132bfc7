#!/usr/bin/env python
"""MGP_WI_DEBUG switch sweep of lap_spmm_wi_kernel on the cfg-C graph in ONE process (development tool).
    python profiles/bench_spmm_modes.py wi|wp m1,m2,...  [ring_kb]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import manifold_gp_b200 as mgp
from manifold_gp_b200 import graph
from manifold_gp_b200.utils import synthetic
kerns = sys.argv[1].split(",")
modes = [int(m) for m in sys.argv[2].split(",")]
rings = [int(r) for r in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0]
n = 1_000_000
dev = torch.device("cuda:0")
x = synthetic.torus(n, seed=0, device=dev)
idx, val = mgp.NearestNeighbors(x).graph(32)
lap = mgp.GraphLaplacianOperator(val, idx, n, torch.tensor([[0.027417]], device=dev), "symmetric", True)
prec = mgp.PrecisionMaternOperator(lap, 2, torch.tensor([[0.5]], device=dev))
st = lap.structure
_, _, diag, a = lap._values()
shift = prec._shift()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
P = torch.randn(n, 16, device=dev); V = torch.empty_like(P)
res = {}
for ring in rings:
    os.environ["MGP_WI_RING_KB"] = str(ring)
    for kern in kerns:
        for m in modes:
            os.environ["MGP_WI_DEBUG"] = str(m)
            graph.SPMM_KERNEL = kern
            for _ in range(4):
                graph.lap_spmm(st, a, diag, P, shift=shift, out=V)
            ev0.record()
            for _ in range(30):
                graph.lap_spmm(st, a, diag, P, shift=shift, out=V)
            ev1.record(); torch.cuda.synchronize()
            graph.SPMM_KERNEL = "auto"
            res[f"{kern}_ring{ring}_debug{m}"] = round(ev0.elapsed_time(ev1) * 1e3 / 30, 1)
print(json.dumps(res))
