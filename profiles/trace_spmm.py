#!/usr/bin/env python
"""NEEDS THE INSTRUMENTED KERNEL of the commit "Temporary: per-tile clock64 trace ..." (reverted right after: the stamps cost ~4 us per
launch); kept with its output (r02_spmm_trace_block0_*.json) as the record of how the numbers were taken.
Per-tile timeline of block 0 of lap_spmm_wi_kernel at cfg-C (MGP_WI_TRACE=1: clock64() stamps of the filler, one helper warp and
two consumer warps per tile; development aid).   python profiles/trace_spmm.py wp|wi"""
import ctypes, json, os, sys
os.environ["MGP_WI_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import manifold_gp_b200 as mgp
from manifold_gp_b200 import graph, _lib
from manifold_gp_b200.utils import synthetic
kern = sys.argv[1] if len(sys.argv) > 1 else "wp"
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
graph.SPMM_KERNEL = kern
for _ in range(5):
    graph.lap_spmm(st, a, diag, P, shift=shift, out=V)
torch.cuda.synchronize()
buf = (ctypes.c_uint64 * (64 * 16))()
_lib._dll.mgp_wi_trace_dump(buf)
t = np.frombuffer(buf, dtype=np.uint64).reshape(64, 16).astype(np.int64)
nt = int((t[:, 0] > 0).sum())
t0 = t[0, 0]
names = ["desc", "region_free", "ids", "copies_issued", "helper_go", "helper_done", "c0_wait", "c0_full", "c0_walked", "c0_release",
         "c15_wait", "c15_full", "c15_walked", "c15_release"]
rows = []
for i in range(nt):
    rows.append({nm: int(t[i, j] - t0) for j, nm in enumerate(names)})
mid = rows[8:nt - 4]
def avg(f):
    return round(float(np.mean([f(r) for r in mid])), 1)
def avgn(f, lag):
    v = [f(rows[i], rows[i + lag]) for i in range(8, nt - 4 - lag)]
    return round(float(np.mean(v)), 1)
summ = {
    "kernel": kern, "tiles_block0": nt, "cycles_total": int(max(t[nt - 1, 9], t[nt - 1, 13]) - t0),
    "cycles_per_tile": avgn(lambda a_, b_: b_["c0_release"] - a_["c0_release"], 1),
    "filler: desc -> region free (waiting for consumers)": avg(lambda r: r["region_free"] - r["desc"]),
    "filler: region free -> ids": avg(lambda r: r["ids"] - r["region_free"]),
    "filler: ids -> copies issued": avg(lambda r: r["copies_issued"] - r["ids"]),
    "filler: copies issued(i) -> desc(i+1)": avgn(lambda a_, b_: b_["desc"] - a_["copies_issued"], 1),
    "helper: region free -> go": avg(lambda r: r["helper_go"] - r["region_free"]),
    "helper: go -> done": avg(lambda r: r["helper_done"] - r["helper_go"]),
    "fill latency: region free -> consumer 0 sees full": avg(lambda r: r["c0_full"] - r["region_free"]),
    "fill latency: helper done -> consumer 0 sees full": avg(lambda r: r["c0_full"] - r["helper_done"]),
    "consumer 0: waiting for full": avg(lambda r: r["c0_full"] - r["c0_wait"]),
    "consumer 0: walk": avg(lambda r: r["c0_walked"] - r["c0_full"]),
    "consumer 0: reduce + epilogue": avg(lambda r: r["c0_release"] - r["c0_walked"]),
    "consumer 0: release(i) -> wait(i+1)": avgn(lambda a_, b_: b_["c0_wait"] - a_["c0_release"], 1),
    "consumer 15: waiting for full": avg(lambda r: r["c15_full"] - r["c15_wait"]),
    "consumer 15: walk": avg(lambda r: r["c15_walked"] - r["c15_full"]),
    "consumer 15: reduce + epilogue": avg(lambda r: r["c15_release"] - r["c15_walked"]),
    "tiles in flight when filled (i - last released)": avg(lambda r: 0),
    "release(i) of consumer 0 -> region_free(i+k) first k with region_free later": None,
}
# how many tiles ahead is the filler: for each tile i, number of tiles j > i whose copies were issued before consumer 0 released i
ahead = []
for i in range(8, nt - 8):
    ahead.append(sum(1 for j in range(i + 1, nt) if rows[j]["copies_issued"] < rows[i]["c0_release"]))
summ["tiles issued ahead of the one being released"] = round(float(np.mean(ahead)), 2)
print(json.dumps(summ, indent=1))
print(json.dumps(rows[20:26]))
