#!/usr/bin/env python
"""Same-process A/B of the two walks of lap_spmm_wi_kernel on the cfg-C graph (N = 1M torus, k = 32, 16 fp32 right-hand sides):
"wi" = one row per 4-lane slot, "wp" = the paired-row walk (union list of two spatially adjacent rows per 8-lane slot).
Interleaved repetitions, CUDA-event timed, inputs larger than L2 rotate between launches.  Also the MGP_WI_DEBUG switches
(1 = no row walk, 2 = no halo copies) for both.  Development tool; bench.py is the judged benchmark.
    python profiles/bench_spmm_pair.py [n] [k]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import manifold_gp_b200 as mgp
from manifold_gp_b200 import graph
from manifold_gp_b200.utils import synthetic

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 32
dev = torch.device("cuda:0")
x = synthetic.torus(n, seed=0, device=dev)
idx, val = mgp.NearestNeighbors(x).graph(k)
lap = mgp.GraphLaplacianOperator(val, idx, n, torch.tensor([[0.027417]], device=dev), "symmetric", True)
prec = mgp.PrecisionMaternOperator(lap, 2, torch.tensor([[0.5]], device=dev))
st = lap.structure
_, _, diag, a = lap._values()
nnz = st.nnz
shift = prec._shift()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
import time
t0 = time.time(); q = st.pair_tiles(); torch.cuda.synchronize(); t_pair = time.time() - t0
t = st.tiles
res = {"n": n, "k": k, "nnz": nnz, "pair_streams_build_s": round(t_pair, 3),
       "tiles": {key: t[key] for key in ("lmax", "hmax", "wnzmax", "nnzw", "qnzmax", "nnzq", "q_unions") if key in t}}
res["slots_per_nnz"] = {"wi": round(t["nnzw"] / nnz, 4), "wp": round(t["nnzq"] / nnz, 4)}
res["stream_bytes_per_launch_MB"] = {"wi": round(t["nnzw"] * 6 / 1e6, 1), "wp": round(t["nnzq"] * 10 / 1e6, 1)}
NB = int(os.environ.get("NB", "3"))        # rotate input/output pairs: 3 x 128 MB > L2
Ps = [torch.randn(n, 16, device=dev) for _ in range(NB)]
Vs = [torch.empty(n, 16, device=dev) for _ in range(NB)]


def timeit(kern, reps=30, warm=4, **kw):
    graph.SPMM_KERNEL = kern
    try:
        for i in range(warm):
            graph.lap_spmm(st, a, diag, Ps[i % NB], shift=shift, out=Vs[i % NB], **kw)
        ev0.record()
        for i in range(reps):
            graph.lap_spmm(st, a, diag, Ps[i % NB], shift=shift, out=Vs[i % NB], **kw)
        ev1.record()
        torch.cuda.synchronize()
    finally:
        graph.SPMM_KERNEL = "auto"
    return round(ev0.elapsed_time(ev1) * 1e3 / reps, 1)


# agreement first
graph.SPMM_KERNEL = "wi"; Ywi = graph.lap_spmm(st, a, diag, Ps[0], shift=shift).clone()
graph.SPMM_KERNEL = "wp"; Ywp = graph.lap_spmm(st, a, diag, Ps[0], shift=shift).clone()
graph.SPMM_KERNEL = "auto"
res["rel_diff_wp_vs_wi"] = float((Ywp.double() - Ywi.double()).norm() / Ywi.double().norm())
alg = nnz * 8 + n * (2 * 16 * 4 + 4)
for rep in range(3):
    for kern in ("wi", "wp"):
        us = timeit(kern)
        res[f"{kern}_rep{rep}"] = {"us": us, "alg_GBs": round(alg / us / 1e3, 0)}
dot = torch.zeros(16, device=dev)
for kern in ("wi", "wp"):
    res[f"{kern}_with_dot"] = timeit(kern, dot_with=Ps[0], dot_out=dot)
for m in (1, 2, 3):
    os.environ["MGP_WI_DEBUG"] = str(m)
    for kern in ("wi", "wp"):
        res[f"{kern}_debug{m}"] = timeit(kern)
os.environ["MGP_WI_DEBUG"] = "0"
print(json.dumps(res))
