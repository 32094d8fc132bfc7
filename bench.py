#!/usr/bin/env python
"""bench.py -- headline benchmark of the IMGP hot path on B200 (BASELINE.json configs[2], "cfg-C"):

    synthetic torus in R^3, N = 1M points, k = 32 (31 out-edges per node), symmetric Laplacian with self loops,
    Matern precision (2nu/kappa^2 + L)^nu with nu = 2, one batched CG solve (16 right-hand sides) to 1e-6.

One "step" = one full CG solve.  Metric = CG iterations per second of that solve (16 right-hand sides advance together;
higher is better) -- a RATE, so that the CPU arm can measure it on a bounded sample of iterations of the same solve instead
of extrapolating a solve it cannot finish (round-1 verdict); the solve time itself is `solve_ms` / `ms_per_step`.  The JSON
line also carries the SpMM roofline (algorithmic bytes of SURVEY.md 8(d) / CUDA-event time of the SpMM launches), the parity
of the timed fp32 solution against an fp64 solve on the same GPU, the kNN build time, the end-to-end number through the
public API with host buffers, and a CPU baseline (oracle port, bounded sample).

    python bench.py                       # N=1 GPU, defaults
    python bench.py --impl reference      # the reference's CPU path (oracle port) on the host cores
    torchrun ... bench.py --gpus 8        # row-partitioned strong scaling (see manifold_gp_b200/distributed.py)
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

# ---- workload definition (every number below is reported in the JSON line) -----------------------
CFG = dict(workload="torus_R3_N1M_k32_nu2_cg16rhs", n=1_000_000, k=32, nu=2, kappa=0.5, rhs=16, tol=1e-6,
           max_iter=4000, normalization="symmetric", self_loops=True, seed=0, rhs_seed=1)
METRIC = "precision_cg_iterations_per_s"
UNIT = "CG iterations/s (N=1M, k=32, nu=2, 16 RHS per iteration)"
CPU_SAMPLE_ITERS = 5          # CG iterations the CPU arms time per step (~2 s each on the box's host cores)


def ncu_traffic(kernel_substr):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the kernel from the round's committed `ncu --set full` capture
    (profiles/r02_ncu_traffic.json, written by profiles/parse_ncu.py from the .ncu-rep of the same build); None when the file has
    no entry for the kernel that actually ran -- never a stale constant."""
    try:
        tab = json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")))
        # graph.LAST_SPMM_KERNEL -> the instantiation ncu lists (template arguments: value type, producer warps, paired walk)
        kernel_substr = {"lap_spmm_wi_kernel<pair>": "lap_spmm_wi_kernel<float, 16, 1>",
                         "lap_spmm_wi_kernel": "lap_spmm_wi_kernel<float, 16, 0>"}.get(kernel_substr, kernel_substr)
        for name, row in tab.items():
            if kernel_substr in name:
                return int(row["dram_bytes"]), row.get("source")
    except Exception:
        pass
    return None, None


def spmm_algorithmic_bytes(n, nnz, c, w=4):
    """nnz*(4 + w) + N*(2*C*w + w)."""
    return nnz * (4 + w) + n * (2 * c * w + w)


class ClockSampler:
    """SM clocks / throttle reasons DURING the timed region (B200_PROFILING.md).  Sampled in-process through NVML (pynvml)
    every 100 ms: polling with an `nvidia-smi -lms` child process was measured to perturb the host-in-the-loop CG solve on
    some boxes (same code 615 ms without it, 700-1150 ms with it); NVML queries from a thread do not.  Falls back to one
    nvidia-smi snapshot at the end of the region if pynvml is unavailable."""

    def __init__(self, index=0):
        self.rows = []          # (sm_mhz, sm_max_mhz, set(reasons))
        self.index = index
        self._stop = threading.Event()
        self._thread = None
        self._nvml = None
        self.period = float(os.environ.get("MGP_BENCH_CLOCK_PERIOD", "0.1"))

    def _loop(self):
        nv, h = self._nvml, self._handle
        names = (("hw_slowdown", getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8)),
                 ("hw_thermal_slowdown", getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40)),
                 ("sw_thermal_slowdown", getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20)),
                 ("sw_power_cap", getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)))
        while not self._stop.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.rows.append((float(sm), float(mx), {n for n, bit in names if mask & bit}))
            except Exception:
                pass
            self._stop.wait(self.period)

    def __enter__(self):
        if self.period <= 0:
            return self
        try:
            import pynvml as nv
            nv.nvmlInit()
            # NVML enumerates physical devices: map the CUDA ordinal through CUDA_VISIBLE_DEVICES when it is a list of indices
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            phys = self.index
            if vis and all(t.strip().isdigit() for t in vis.split(",")):
                phys = int(vis.split(",")[self.index])
            self._handle = nv.nvmlDeviceGetHandleByIndex(phys)
            self._nvml = nv
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()
        except Exception:
            self._nvml = None
        return self

    def __exit__(self, *a):
        if self._thread is not None:
            self._stop.set()
            self._thread.join(timeout=2)
        elif not self.rows:
            q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=10).stdout.strip().split(",")
                names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
                self.rows.append((float(out[0]), float(out[1]),
                                  {n for n, v in zip(names, out[2:6]) if v.strip().lower().startswith("active")}))
            except Exception:
                pass

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(r[0] for r in self.rows)
        reasons = set()
        for r in self.rows:
            reasons |= r[2]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(r[1] for r in self.rows), "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvml" if self._nvml is not None else "nvidia-smi"}


def dist_info():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# =====================================================================================================================
# our arm
# =====================================================================================================================
def build_problem(dev, n, k, seed):
    import manifold_gp_b200 as mgp
    from manifold_gp_b200.utils import synthetic
    x = synthetic.torus(n, seed=seed, device=dev)
    knn = mgp.NearestNeighbors(x)
    knn.search(x[:4096].contiguous(), k)              # warm-up: module load + kernel attributes, not the search
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    dist, nbr = knn.search(x, k)
    torch.cuda.synchronize()
    t_search = time.perf_counter() - t0
    t1 = time.perf_counter()
    idx, val = knn.graph(k)
    torch.cuda.synchronize()
    t_graph = time.perf_counter() - t1                # search again + symmetrise (graph() searches itself)
    eps = float(dist[:, k - 1].sqrt().median())       # graph bandwidth = median k-th-NN distance
    return x, idx, val, eps, t_search, t_graph


def run_ours(args):
    import manifold_gp_b200 as mgp
    from manifold_gp_b200 import graph, solvers, _lib
    rank, world, local = dist_info()
    if world > 1:
        from manifold_gp_b200 import distributed
        return distributed.bench_main(args, CFG, clock_sampler=ClockSampler)
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    n, k, c = args.n, CFG["k"], CFG["rhs"]
    x, idx, val, eps, t_search, t_graph = build_problem(dev, n, k, CFG["seed"])
    m = idx.shape[1]
    nnz = 2 * m
    t0 = time.perf_counter()
    lap = mgp.GraphLaplacianOperator(val, idx, n, torch.tensor([[eps]], device=dev), CFG["normalization"], CFG["self_loops"])
    prec = mgp.PrecisionMaternOperator(lap, CFG["nu"], torch.tensor([[CFG["kappa"]]], device=dev))
    lap.structure
    torch.cuda.synchronize()
    t_struct = time.perf_counter() - t0
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(); lap._values(); ev1.record(); torch.cuda.synchronize()
    t_values_ms = ev0.elapsed_time(ev1)

    g = torch.Generator(device=dev).manual_seed(CFG["rhs_seed"])
    B = torch.randn(n, c, device=dev, generator=g)
    Bh = B.cpu().pin_memory()
    Xh = torch.empty_like(Bh).pin_memory()

    def solve_dev():
        return solvers.linear_cg(prec, B, tolerance=CFG["tol"], max_iter=CFG["max_iter"], return_info=True)

    def solve_e2e():
        b = Bh.to(dev, non_blocking=True)
        xs, info = solvers.linear_cg(prec, b, tolerance=CFG["tol"], max_iter=CFG["max_iter"], return_info=True)
        Xh.copy_(xs, non_blocking=True)
        torch.cuda.synchronize()
        return info

    import warnings
    warnings.simplefilter("ignore", RuntimeWarning)
    for _ in range(args.warmup):
        sol, info = solve_dev()
    torch.cuda.synchronize()
    _lib.reset_launch_count()
    with ClockSampler(local) as clk:
        torch.cuda.synchronize()
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
        ev0.record()
        marks[0].record()
        for i in range(args.steps):
            sol, info = solve_dev()
            marks[i + 1].record()
        ev1.record()
        torch.cuda.synchronize()
    per_solve_ms = [round(marks[i].elapsed_time(marks[i + 1]), 1) for i in range(args.steps)]
    launches = _lib.launch_count()
    ms_step = ev0.elapsed_time(ev1) / args.steps
    iters = info["iterations"]

    # ---- parity of the timed result ----------------------------------------------------------------------------------------
    # (1) true residual b - A x of what the timed solve returned (fp64 accumulation of the norms) and of the same solve
    #     without the final true-residual correction (settings.cg_polish) -- the reference's fp32 mBCG returns the latter;
    # (2) the same system solved in fp64 ON THIS GPU (fp64 kernels, masks of the published algorithm pushed out of the way):
    #     per-column relative error of the fp32 solution -- north-star: CG solutions within 1e-4.
    from manifold_gp_b200 import settings
    Bn = B.double().norm(dim=0)
    true_rel_cols = (prec.matmul(sol) - B).double().norm(dim=0) / Bn
    true_rel = float(true_rel_cols.mean())
    polish = dict(info.get("polish") or {})
    with settings.cg_polish(False):
        sol_np, info_np = solve_dev()
    true_rel_unpolished = float(((prec.matmul(sol_np) - B).double().norm(dim=0) / Bn).mean())
    parity = {"fp64_reference": "unavailable"}
    try:
        lap64 = mgp.GraphLaplacianOperator(val.double(), idx, n, torch.tensor([[eps]], device=dev, dtype=torch.float64),
                                           CFG["normalization"], CFG["self_loops"])
        prec64 = mgp.PrecisionMaternOperator(lap64, CFG["nu"], torch.tensor([[CFG["kappa"]]], device=dev, dtype=torch.float64))
        t64 = time.perf_counter()
        x64, info64 = solvers.linear_cg(prec64, B.double(), tolerance=1e-10, eps=1e-30, stop_updating_after=1e-30,
                                        max_iter=3 * CFG["max_iter"], return_info=True)
        torch.cuda.synchronize()
        t64 = time.perf_counter() - t64
        res64 = float(((prec64.matmul(x64) - B.double()).norm(dim=0) / Bn).max())
        xn = x64.norm(dim=0)
        err = (sol.double() - x64).norm(dim=0) / xn
        err_np = (sol_np.double() - x64).norm(dim=0) / xn
        parity = {"fp64_reference": "same system solved with the fp64 kernels on this GPU (tolerance 1e-10, eps masks disabled)",
                  "fp64_iterations": info64["iterations"], "fp64_true_relative_residual_max": res64, "fp64_solve_s": round(t64, 3),
                  "solution_rel_err_max": float(err.max()), "solution_rel_err_mean": float(err.mean()),
                  "solution_rel_err_unpolished_max": float(err_np.max()), "within_1e-4": bool(float(err.max()) <= 1e-4)}
        del lap64, prec64, x64
    except Exception as e:      # reported, never hidden
        parity = {"fp64_reference": "failed: " + repr(e)[:200]}
    del sol_np
    torch.cuda.empty_cache()

    # ---- end-to-end through the public API with host buffers ---------------------------------------------------------
    solve_e2e()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(max(1, min(args.steps, 3))):
        solve_e2e()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / max(1, min(args.steps, 3))

    # ---- dominant kernel: the C=16 SpMM, timed with CUDA events over back-to-back launches (inputs >> L2) --------------
    _, _, diag, a = lap._values()
    shift = prec._shift()
    P = torch.randn(n, c, device=dev)
    V = torch.empty_like(P)
    reps = 50
    for _ in range(5):
        graph.lap_spmm(lap.structure, a, diag, P, shift=shift, out=V)
    ev0.record()
    for _ in range(reps // 2):
        graph.lap_spmm(lap.structure, a, diag, P, shift=shift, out=V)
        graph.lap_spmm(lap.structure, a, diag, V, shift=shift, out=P)
    ev1.record(); torch.cuda.synchronize()
    spmm16_us = ev0.elapsed_time(ev1) * 1e3 / (2 * (reps // 2))
    spmm16_kernel = graph.LAST_SPMM_KERNEL
    p1 = torch.randn(n, 1, device=dev); v1 = torch.empty_like(p1)
    for _ in range(5):
        graph.lap_spmm(lap.structure, a, diag, p1, out=v1)
    ev0.record()
    for _ in range(reps // 2):
        graph.lap_spmm(lap.structure, a, diag, p1, out=v1)
        graph.lap_spmm(lap.structure, a, diag, v1, out=p1)
    ev1.record(); torch.cuda.synchronize()
    spmv1_us = ev0.elapsed_time(ev1) * 1e3 / (2 * (reps // 2))
    spmv1_kernel = graph.LAST_SPMM_KERNEL

    knn_tensor = knn_tensor_bench(dev)

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    b16 = spmm_algorithmic_bytes(n, nnz, c)
    b1 = spmm_algorithmic_bytes(n, nnz, 1)
    ach16 = b16 / (spmm16_us * 1e-6) / 1e9
    ach1 = b1 / (spmv1_us * 1e-6) / 1e9

    total_iters = iters + int(polish.get("iterations", 0))
    traffic, traffic_src = ncu_traffic(spmm16_kernel)
    out = {
        "metric": METRIC, "value": round(total_iters / (ms_step * 1e-3), 1), "unit": UNIT, "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms_step, 3), "solve_ms": round(ms_step, 3), "per_step_ms": per_solve_ms,
        "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": CFG["workload"], "n": n, "k": k, "edges_M": m, "nnz": nnz, "nu": CFG["nu"], "kappa": CFG["kappa"],
                   "eps": round(eps, 6), "rhs": c, "tol": CFG["tol"], "normalization": CFG["normalization"],
                   "self_loops": CFG["self_loops"], "l2": "inputs larger than L2 (matrix 8*nnz B + 4 vectors of N*16*4 B >> 126 MB)"},
        "cg_iterations": iters, "cg_polish_iterations": int(polish.get("iterations", 0)),
        "cg_converged": bool(info["converged"]) and true_rel <= 10 * CFG["tol"],
        "cg_converged_recurrence": bool(info["converged"]), "cg_recurrence_residual": info["mean_residual"],
        "cg_true_relative_residual": true_rel, "cg_true_relative_residual_max_col": float(true_rel_cols.max()),
        "cg_true_relative_residual_unpolished": true_rel_unpolished,
        "cg_note": "linear_cg (and the reference's fp32 mBCG) stops on the recurrence residual; `unpolished` is what that returns. "
                   "The timed solve includes settings.cg_polish (one true-residual correction, `cg_polish_iterations` extra iterations); "
                   "`cg_converged` requires the TRUE residual within 10x of tol.",
        "parity": parity,
        "e2e": {"value": round(total_iters / (e2e_ms * 1e-3), 1), "unit": UNIT, "solve_ms": round(e2e_ms, 3),
                "h2d_bytes_per_step": Bh.numel() * 4, "d2h_bytes_per_step": Xh.numel() * 4},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": f"{spmm16_kernel} fp32 (C=16 SpMM step of the Matern precision operator)",
                     "achieved": round(ach16, 1), "peak": hbm_peak, "unit": "GB/s", "frac": round(ach16 / hbm_peak, 4),
                     "frac_of_nominal_8000": round(ach16 / 8000.0, 4), "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": b16, "us_per_launch": round(spmm16_us, 2),
                     "spmm_share_of_step": round(2 * iters * spmm16_us * 1e-3 / ms_step, 3)},
        "spmv_c1": {"kernel": f"{spmv1_kernel}<float> (one right-hand side: the plain Laplacian SpMV)", "us_per_launch": round(spmv1_us, 2),
                    "frac_of_nominal_8000": round(ach1 / 8000.0, 4),
                    "achieved_gbs": round(ach1, 1), "frac": round(ach1 / hbm_peak, 4),
                    "traffic": ncu_traffic(spmv1_kernel)[0],
                    "algorithmic_bytes_per_launch": b1},
        "knn_build_s": round(t_search, 4), "knn_build_candidates_per_s": round(float(n) * float(n) / max(t_search, 1e-9), 1),
        "graph_symmetrize_s": round(max(t_graph - t_search, 0.0), 4),
        "structure_build_s": round(t_struct, 4), "laplacian_values_ms": round(t_values_ms, 3),
        "knn_tensor": knn_tensor,
        "clocks": clk.summary(),
    }
    if not args.no_cfgb:
        out["cfgB"] = cfgb_bench()
    if not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(idx.cpu(), val.cpu(), n, eps)
    print(json.dumps(out))
    return out


def cfgb_bench():
    """BASELINE.json configs[1] end to end in a child process (profiles/run_cfgB.py: 70k x 784 cloud, k = 10 -> tcgen05 kNN graph
    -> smallest-500 eigenpairs -> spectral features / out-of-sample extension -> semi-supervised posterior -> one 16-RHS
    precision CG solve), stage timings and self-checks as that script reports them."""
    try:
        r = subprocess.run([sys.executable, os.path.join(ROOT, "profiles", "run_cfgB.py")], capture_output=True, text=True, timeout=600)
        line = [l for l in r.stdout.splitlines() if l.startswith("{")]
        if r.returncode != 0 or not line:
            return {"error": (r.stderr or r.stdout)[-300:]}
        return json.loads(line[-1])
    except Exception as e:      # reported, never hidden
        return {"error": repr(e)[:300]}


def knn_tensor_bench(dev, n=70000, d=784, k=10):
    """BASELINE cfg-B's kNN (RMNIST-shape cloud) through the tcgen05 search, against the tensor-pipe roofline: the
    denominator is cuBLAS TF32 (torch.matmul, 8192^3) measured in this process, as SURVEY.md 8(d) prescribes when TF32 MMA
    is the instruction used; `issued` counts the 3 TF32 products per element pair, `useful` the 2*Q*N*d of the problem."""
    import manifold_gp_b200 as mgp
    from manifold_gp_b200.utils import synthetic
    try:
        x = synthetic.rmnist_shape(n, d, device=dev)
        knn = mgp.NearestNeighbors(x)
        knn.search(x, k)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        e0.record()
        for _ in range(reps):
            knn.search(x, k)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        info = knn.last_search
        research = int(info["stats"][0]) if "stats" in info else None
        old = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = True
        a = torch.randn(8192, 8192, device=dev); b = torch.randn(8192, 8192, device=dev)
        for _ in range(3):
            a @ b
        e0.record()
        for _ in range(10):
            a @ b
        e1.record()
        torch.cuda.synchronize()
        torch.backends.cuda.matmul.allow_tf32 = old
        peak = 2.0 * 8192 ** 3 / (e0.elapsed_time(e1) / 10) / 1e9
        del a, b, x
        flops = 2.0 * n * n * d
        return {"workload": f"rmnist_shape_{n}x{d}_k{k}", "kernel": info["kernel"], "ms": round(ms, 3),
                "useful_tflops": round(flops / ms / 1e9, 1), "issued_tf32_tflops": round(3 * flops / ms / 1e9, 1),
                "peak_tf32_tflops_cublas_measured": round(peak, 1), "frac_issued_of_peak": round(3 * flops / ms / 1e9 / peak, 4),
                "research_queries": research, "timed": "whole search (prep + sweep + re-rank + re-search), CUDA events"}
    except Exception as e:      # reported, never hidden
        return {"error": repr(e)[:300]}


# =====================================================================================================================
# CPU arms: the reference's torch-sparse path restated (oracle), timed on the host cores
# =====================================================================================================================
def cpu_baseline(idx, val, n, eps, sample_iters=CPU_SAMPLE_ITERS, threads=None, state=None):
    """Time `sample_iters` consecutive CG iterations of the same solve with the oracle port (index_select -> mul -> scatter_add
    on the int64 upper-triangular COO, exactly what torch_sparse.spmm lowers to) and report the measured RATE -- nothing is
    extrapolated.  `state` carries (operator, x, r, p) from one call to the next so successive steps continue the same solve."""
    import oracle
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    if state is None or "A" not in state:
        olap = oracle.LaplacianOracle(val, idx, n, eps, CFG["normalization"], CFG["self_loops"])
        olap.laplacian_triu  # value build (untimed here)
        A = lambda v: oracle.precision_matmul(olap, CFG["nu"], CFG["kappa"], v)
        g = torch.Generator().manual_seed(CFG["rhs_seed"])
        b = torch.randn(n, CFG["rhs"], generator=g)
        b = b / b.norm(dim=0)
        A(b[:, :1])  # warm-up
        st = {"A": A, "x": torch.zeros_like(b), "r": b.clone(), "p": b.clone()}
        if state is not None:
            state.update(st)
        else:
            state = st
    A, x, r, p = state["A"], state["x"], state["r"], state["p"]
    t0 = time.perf_counter()
    for _ in range(sample_iters):
        v = A(p)
        rz = (r * r).sum(0)
        alpha = rz / (p * v).sum(0)
        x = x + alpha * p
        r = r - alpha * v
        beta = (r * r).sum(0) / rz
        p = r + beta * p
    dt = time.perf_counter() - t0
    state.update(x=x, r=r, p=p)
    return {"value": round(sample_iters / dt, 4), "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{sample_iters} consecutive CG iterations of the same N={n} solve ({CFG['nu']} Laplacian matvecs with "
                      f"C={CFG['rhs']} each per iteration) in {dt:.2f} s; rate measured, nothing extrapolated",
            "sample_iterations": sample_iters, "sample_seconds": round(dt, 3), "cpu_model": _cpu_model()}


def _cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def run_reference(args):
    """The reference's own CPU path (oracle port; the reference itself cannot be imported: gpytorch / linear_operator /
    torch_sparse / faiss are absent and there is no network).  Graph build (untimed setup) uses scipy's kd-tree."""
    rank, world, _ = dist_info()
    if rank != 0:
        return None
    import numpy as np
    import oracle
    from scipy.spatial import cKDTree
    threads = os.cpu_count()
    torch.set_num_threads(threads)
    n, k = args.n, CFG["k"]
    x = oracle.datasets.torus(n, seed=CFG["seed"])
    tree = cKDTree(x.numpy())
    d, i = tree.query(x.numpy(), k=k, workers=-1)
    d2 = torch.from_numpy((d.astype(np.float32)) ** 2)
    eps = float(np.median(d[:, k - 1]))
    idx, val = oracle.symmetrize_coalesce(d2, torch.from_numpy(i.astype(np.int64)), n)
    rates, secs = [], []
    base, state = None, {}
    for s_ in range(args.warmup + args.steps):
        base = cpu_baseline(idx, val, n, eps, sample_iters=CPU_SAMPLE_ITERS, threads=threads, state=state)
        if s_ >= args.warmup:
            rates.append(base["value"]); secs.append(base["sample_seconds"])
    rate = sum(rates) / len(rates)
    base["value"] = round(rate, 4)
    ms = 1e3 * sum(secs) / len(secs)
    out = {"impl": "reference", "metric": METRIC, "value": round(rate, 4), "unit": UNIT, "n_gpus": 0,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 1), "higher_is_better": True,
           "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": CFG["workload"], "n": n, "k": k, "edges_M": int(idx.shape[1]), "nnz": 2 * int(idx.shape[1]),
                      "nu": CFG["nu"], "kappa": CFG["kappa"], "eps": round(eps, 6), "rhs": CFG["rhs"], "tol": CFG["tol"],
                      "normalization": CFG["normalization"], "self_loops": CFG["self_loops"],
                      "l2": "inputs larger than L2 (matrix 8*nnz B + 4 vectors of N*16*4 B >> 126 MB)"},
           "step": f"one step = {CPU_SAMPLE_ITERS} consecutive CG iterations of the solve (a bounded sample; the full solve needs "
                   "~1700 iterations = ~1 h on these cores)",
           "cpu_baseline": base,
           "e2e": {"value": round(rate, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=CFG["n"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cfgb", action="store_true", help="skip the cfg-B end-to-end child run")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
