#!/usr/bin/env python
"""Profiling driver: cfg-C problem, a short CG run (fixed iteration count) and a few standalone SpMM launches.
Used under ncu (launch list + one --set full capture); never a source of bench numbers."""
import argparse
import os
import sys
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--rhs", type=int, default=16)
    args = ap.parse_args()
    import manifold_gp_b200 as mgp
    from manifold_gp_b200 import graph, solvers
    from manifold_gp_b200.utils import synthetic
    dev = torch.device("cuda:0")
    x = synthetic.torus(args.n, seed=0, device=dev)
    knn = mgp.NearestNeighbors(x)
    idx, val = knn.graph(32)
    lap = mgp.GraphLaplacianOperator(val, idx, args.n, torch.tensor([[0.027417]], device=dev), "symmetric", True)
    prec = mgp.PrecisionMaternOperator(lap, 2, torch.tensor([[0.5]], device=dev))
    B = torch.randn(args.n, args.rhs, device=dev)
    warnings.simplefilter("ignore")
    with mgp.settings.max_lanczos_quadrature_iterations(min(20, args.iters)):
        solvers.linear_cg(prec, B, tolerance=0.0, max_iter=args.iters)
    _, _, diag, a = lap._values()
    p1 = torch.randn(args.n, 1, device=dev)
    v1 = torch.empty_like(p1)
    for _ in range(4):
        graph.lap_spmm(lap.structure, a, diag, p1, out=v1)
    torch.cuda.synchronize()
    print("profile run ok")


if __name__ == "__main__":
    main()
