#!/usr/bin/env python
"""BASELINE.json configs[4] ("cfg-E"): kNN graph build + Laplacian SpMV sweep over N, d, k on one B200, with the reference's
CPU torch-sparse path (oracle port) timed on the box's host cores for the points it can finish in seconds.

    python profiles/run_cfgE_sweep.py [--quick] > profiles/r02_cfgE_sweep.json

Per point: kNN search (kernel used, seconds, candidates/s, useful TFLOP/s), symmetrise + structure build, Laplacian value
build, the C = 1 SpMV and the C = 16 SpMM (us, algorithmic GB/s of SURVEY.md 8d, fraction of the measured HBM peak), and the CPU
matvec.  N = 100M is out of reach of an exhaustive search inside the round's GPU budget (10^16 candidates = ~70 GPU-minutes at
the measured 2.4e12 candidates/s); the sweep states that instead of extrapolating."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import manifold_gp_b200 as mgp
from manifold_gp_b200 import graph
from manifold_gp_b200.utils import synthetic

quick = "--quick" in sys.argv
dev = torch.device("cuda:0")
peak = 6543.4
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def gpu_time(fn, reps=1, warm=0):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(reps):
        out = fn()
    ev1.record()
    torch.cuda.synchronize()
    return ev0.elapsed_time(ev1) * 1e-3 / reps, out


def cpu_matvec_ms(idx, val, n, eps, c):
    import oracle
    torch.set_num_threads(os.cpu_count())
    olap = oracle.LaplacianOracle(val.cpu(), idx.cpu(), n, eps, "symmetric", True)
    v = torch.randn(n, c)
    olap.matmul(v)
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); olap.matmul(v); ts.append(time.perf_counter() - t0)
    return sorted(ts)[1] * 1e3


grid = []
for n in (100_000, 1_000_000):
    for d in (3, 64, 784):
        for k in (8, 32, 64):
            grid.append((n, d, k))
grid += [(10_000_000, 3, 8), (10_000_000, 3, 32)]
if quick:
    grid = [(100_000, 3, 8), (100_000, 64, 32), (200_000, 3, 64)]
rows = []
cache = {}
for n, d, k in grid:
    row = {"n": n, "d": d, "k": k}
    try:
        if (n, d) not in cache:
            cache.clear()
            cache[(n, d)] = synthetic.torus(n, seed=0, device=dev) if d == 3 else synthetic.rmnist_shape(n, d, device=dev)
        x = cache[(n, d)]
        knn = mgp.NearestNeighbors(x)
        knn.search(x[:2048].contiguous(), k)                      # module / attribute warm-up
        if n <= 1_000_000:
            t_s, (d2, _) = gpu_time(lambda: knn.search(x, k))
            row.update(knn_search_s=round(t_s, 4), knn_kernel=knn.last_search["kernel"],
                       knn_candidates_per_s=round(float(n) * n / t_s, 1), knn_useful_tflops=round(2.0 * n * n * d / t_s / 1e12, 2))
            eps = float(d2[:, k - 1].sqrt().median())
            del d2
        t_g, (idx, val) = gpu_time(lambda: knn.graph(k))            # search + symmetrise / coalesce (the reference's graph())
        row.update(graph_s=round(t_g, 4), knn_kernel=knn.last_search["kernel"], edges_M=int(idx.shape[1]))
        if n > 1_000_000:
            dk, _ = knn.search(x[:65536].contiguous(), k)
            eps = float(dk[:, k - 1].sqrt().median())
            row.update(knn_search_s_upper_bound=round(t_g, 4), knn_candidates_per_s_lower_bound=round(float(n) * n / t_g, 1))
        m = int(idx.shape[1]); nnz = 2 * m
        lap = mgp.GraphLaplacianOperator(val, idx, n, torch.tensor([[eps]], device=dev), "symmetric", True)
        t_st, _ = gpu_time(lambda: lap.structure)
        t_v, (_, _, diag, a) = gpu_time(lambda: lap._values())
        row.update(structure_s=round(t_st, 4), values_ms=round(t_v * 1e3, 3), eps=round(eps, 6))
        st = lap.structure
        for c in (1, 16):
            P = torch.randn(n, c, device=dev); V = torch.empty_like(P)
            def two():
                graph.lap_spmm(st, a, diag, P, out=V); graph.lap_spmm(st, a, diag, V, out=P)
            t, _ = gpu_time(two, reps=10, warm=2)
            us = t * 1e6 / 2
            alg = nnz * 8 + n * (2 * c * 4 + 4)
            row[f"spmm_c{c}"] = {"kernel": graph.LAST_SPMM_KERNEL, "us": round(us, 1), "alg_GBs": round(alg / us / 1e3, 1),
                                 "frac_of_measured_hbm_peak": round(alg / us / 1e3 / peak, 3)}
            del P, V
        if n <= 1_000_000 and not (d == 784 and k != 32):
            for c in (1, 16):
                row[f"cpu_matvec_c{c}_ms"] = round(cpu_matvec_ms(idx, val, n, eps, c), 1)
            row["cpu_cores"] = os.cpu_count()
            row["speedup_c1"] = round(row["cpu_matvec_c1_ms"] * 1e3 / row["spmm_c1"]["us"], 1)
        del lap, st, idx, val, a, diag, knn
        torch.cuda.empty_cache()
    except Exception as e:      # reported, never hidden
        row["error"] = repr(e)[:300]
    rows.append(row)
    print(json.dumps(row), file=sys.stderr, flush=True)
print(json.dumps({"sweep": "cfg-E kNN build + Laplacian SpMV, one B200", "hbm_peak_gbs": peak,
                  "not_run": "N = 100M: exhaustive search needs ~1e16 candidates (~70 GPU-minutes at 2.4e12/s); d = 784 at N = 10M: 31 GB of points, "
                             "157 PFLOP (~15 min)", "rows": rows}))
