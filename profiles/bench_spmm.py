#!/usr/bin/env python
"""Kernel micro-benchmarks on the cfg-C graph (N=1M torus, k=32): SpMM variants and CG vector kernels, CUDA-event timed,
inputs far larger than L2.  Development tool; bench.py is the judged benchmark."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import manifold_gp_b200 as mgp
from manifold_gp_b200 import graph, solvers, _lib
from manifold_gp_b200.utils import synthetic

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
dev = torch.device("cuda:0")
x = synthetic.torus(n, seed=0, device=dev)
idx, val = mgp.NearestNeighbors(x).graph(32)
lap = mgp.GraphLaplacianOperator(val, idx, n, torch.tensor([[0.027417]], device=dev), "symmetric", True)
prec = mgp.PrecisionMaternOperator(lap, 2, torch.tensor([[0.5]], device=dev))
st = lap.structure
_, _, diag, a = lap._values()
nnz = st.nnz
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def timeit(fn, reps=30, warm=5):
    for _ in range(warm):
        fn()
    ev0.record()
    for _ in range(reps):
        fn()
    ev1.record()
    torch.cuda.synchronize()
    return ev0.elapsed_time(ev1) * 1e3 / reps


ONLY = os.environ.get("BENCH_KERNELS", "").split(",") if os.environ.get("BENCH_KERNELS") else None
t = st.tiles
res = {"tiles": {k: t[k] for k in ("lmax", "nzmax", "pnzmax", "wnzmax", "nnzp", "nnzw", "halo_total") if k in t}, "nnz": nnz}
shift = prec._shift()
MODES = [int(m) for m in os.environ.get("WI_MODES", "").split(",") if m]
if MODES:   # A/B of the v5 kernel's experiment switches inside ONE process (same box, same clocks)
    P = torch.randn(n, 16, device=dev); V = torch.empty_like(P)
    graph.SPMM_KERNEL = "wi"
    for rep in range(2):
        for m in MODES:
            os.environ["MGP_WI_DEBUG"] = str(m)
            res[f"wi_mode{m}_rep{rep}"] = round(timeit(lambda: graph.lap_spmm(st, a, diag, P, shift=shift, out=V)), 1)
    print(json.dumps(res)); sys.exit(0)
for c in (1, 4, 8, 16):
    P = torch.randn(n, c, device=dev); V = torch.empty_like(P)
    dot = torch.zeros(c, device=dev)
    alg = nnz * 8 + n * (2 * c * 4 + 4)
    for kern in ("csr", "tiled", "pipe", "wi"):
        if (kern == "pipe" and c % 4) or (kern == "wi" and c % 16) or (ONLY and kern not in ONLY):
            continue
        graph.SPMM_KERNEL = kern
        t = timeit(lambda: graph.lap_spmm(st, a, diag, P, shift=shift, out=V))
        td = timeit(lambda: graph.lap_spmm(st, a, diag, P, shift=shift, out=V, dot_with=P, dot_out=dot))
        res[f"spmm_{kern}_c{c}"] = {"us": round(t, 1), "us_with_dot": round(td, 1), "alg_GBs": round(alg / t / 1e3, 0)}
    graph.SPMM_KERNEL = "auto"
# CG vector kernels via a short solve timing split
import warnings
warnings.simplefilter("ignore")
if os.environ.get("BENCH_NO_CG"):
    print(json.dumps(res)); sys.exit(0)
B = torch.randn(n, 16, device=dev)
for _ in range(2):
    solvers.linear_cg(prec, B, tolerance=0.0, max_iter=50, max_tridiag_iter=20)
torch.cuda.synchronize()
ev0.record()
_, info = solvers.linear_cg(prec, B, tolerance=0.0, max_iter=200, max_tridiag_iter=20, return_info=True)
ev1.record(); torch.cuda.synchronize()
res["cg_iter_us"] = round(ev0.elapsed_time(ev1) * 1e3 / 200, 1)
print(json.dumps(res))
