"""Timing + diagnostics of the tcgen05 kNN search (knn_tc.cu) vs the CUDA-core kernel.  Usage:
    python profiles/prof_knn_tc.py [n] [d] [k] [--skip-cc]"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import manifold_gp_b200 as mgp  # noqa: E402
from manifold_gp_b200.utils import synthetic  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
n = int(args[0]) if len(args) > 0 else 70000
d = int(args[1]) if len(args) > 1 else 784
k = int(args[2]) if len(args) > 2 else 10
dev = torch.device("cuda:0")
x = synthetic.torus(n, device=dev) if "--torus" in sys.argv else synthetic.rmnist_shape(n, d, device=dev)
d = x.shape[1]
knn = mgp.NearestNeighbors(x)


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2], out


t_tc, (d_tc, i_tc) = timed(lambda: knn.search(x, k))
info = knn.last_search
res = {"n": n, "d": d, "k": k, "kernel": info["kernel"], "tc_ms": t_tc}
if info["kernel"] == "tcgen05":
    st = info["stats"].cpu()
    res.update(research=int(st[0]), max_abs_err=float(st[1:2].view(torch.float32)), processed=int(st[2]))
    flops = 2.0 * n * n * d
    res["useful_tflops"] = flops / t_tc / 1e9
    res["issued_tf32_tflops"] = 3 * flops / t_tc / 1e9
if "--skip-cc" not in sys.argv:
    knn.tensor_core = False
    t_cc, (d_cc, i_cc) = timed(lambda: knn.search(x, k), reps=1)
    res.update(cc_ms=t_cc, idx_equal=bool(torch.equal(i_tc, i_cc)), dist_equal=bool(torch.equal(d_tc, d_cc)),
               rows_differ=int((i_tc != i_cc).any(1).sum()))
print(json.dumps(res))
