#!/usr/bin/env python
"""Roofline numbers of the SECONDARY kernels at cfg-C (N = 1M torus, k = 32): the per-bandwidth value build (3 passes), its paired /
single-row stream layouts, the SDDMM of the backward pass, the Lanczos re-orthogonalisation kernels.  Algorithmic bytes (stated per
kernel below) over CUDA-event time, against MEASURED_PEAKS.json hbm_gbs.  Development tool; bench.py is the judged benchmark.
    python profiles/bench_secondary.py"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import manifold_gp_b200 as mgp
from manifold_gp_b200 import graph, _lib
from manifold_gp_b200._lib import c_int32, c_int64, ptr, stream
from manifold_gp_b200.utils import synthetic

n, k = 1_000_000, 32
dev = torch.device("cuda:0")
peak = 6650.0
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
x = synthetic.torus(n, seed=0, device=dev)
idx, val = mgp.NearestNeighbors(x).graph(k)
lap = mgp.GraphLaplacianOperator(val, idx, n, torch.tensor([[0.027417]], device=dev), "symmetric", True)
st = lap.structure
nnz = st.nnz
d2 = st.d2csr(val)
_, _, diag, a = lap._values()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def timeit(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    ev0.record()
    for _ in range(reps):
        fn()
    ev1.record(); torch.cuda.synchronize()
    return ev0.elapsed_time(ev1) * 1e3 / reps


res = {"n": n, "k": k, "nnz": nnz, "peak_gbs": peak}


def row(name, us, nbytes, note):
    res[name] = {"us": round(us, 1), "algorithmic_MB": round(nbytes / 1e6, 1), "GBs": round(nbytes / us / 1e3, 0),
                 "frac_of_measured_peak": round(nbytes / us / 1e3 / peak, 3), "bytes": note}


eps = torch.tensor([0.027417], device=dev)
us = timeit(lambda: graph.lap_values(st, d2, eps, True))
row("lap_values (3 passes)", us, nnz * (3 * 8 + 4) + n * 4 * 8, "3 x (col 4 B + d2 4 B) read + a 4 B written per entry, 8 degree-vector passes of N x 4 B")
t = st.tiles
us = timeit(lambda: _lib.call("mgp_lap_wi_values_f32", ptr(st.rowptr), ptr(t["wptr"]), ptr(a), c_int64(n), ptr(torch.empty(t["nnzw"] + 64, device=dev)), stream()))
row("lap_wi_values (single-row stream layout)", us, nnz * 4 + t["nnzw"] * 4, "a read + stream written")
tq = st.pair_tiles()
out_q = torch.empty(tq["qsrc"].numel(), device=dev)
us = timeit(lambda: _lib.call("mgp_lap_pair_values_f32", ptr(tq["qsrc"]), ptr(a), c_int64(out_q.numel()), ptr(out_q), stream()))
row("lap_pair_values (paired stream layout)", us, out_q.numel() * 8 + nnz * 4, "source index 4 B read + value 4 B written per slot, a gathered once")
X = torch.randn(n, 16, device=dev); G = torch.randn(n, 16, device=dev)
us = timeit(lambda: graph.lap_sddmm(st, G, X))
row("lap_sddmm (C = 16)", us, nnz * (4 + 4) + 2 * n * 16 * 4 + n * 4, "col read + g_a written per entry, both [N,16] blocks once (the per-entry 64-byte row gathers are L2 traffic)")
for j in (32, 128):
    q = torch.randn(j + 1, n, device=dev); r = torch.randn(n, device=dev)
    c = torch.zeros(j + 1, device=dev); nrm2 = torch.zeros(1, device=dev)
    ws = torch.zeros(_lib.query("mgp_lanczos_ws_bytes", c_int64(n), c_int32(j + 1)), dtype=torch.uint8, device=dev)
    us = timeit(lambda: _lib.call("mgp_lanczos_dots_f32", ptr(q), c_int64(n), c_int32(j), ptr(r), c_int64(n), ptr(c), ptr(ws), stream()))
    row(f"lanczos_dots (j = {j})", us, (j + 1) * n * 4, "Q[0:j] and r read once")
    us = timeit(lambda: _lib.call("mgp_lanczos_axpy_f32", ptr(q), c_int64(n), c_int32(j), ptr(r), c_int64(n), ptr(c), ptr(nrm2), ptr(ws), stream()))
    row(f"lanczos_axpy (j = {j})", us, (j + 2) * n * 4, "Q[0:j] read, r read + written")
    del q
print(json.dumps(res, indent=1))
