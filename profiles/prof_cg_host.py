"""Host-side profile (cProfile) of one full cfg-C CG solve: where does wall time go besides the GPU kernels?"""
import cProfile, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import manifold_gp_b200 as mgp
from manifold_gp_b200 import solvers
from manifold_gp_b200.utils import synthetic
import warnings; warnings.simplefilter("ignore")

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
dev = torch.device("cuda:0")
x = synthetic.torus(n, seed=0, device=dev)
idx, val = mgp.NearestNeighbors(x).graph(32)
lap = mgp.GraphLaplacianOperator(val, idx, n, torch.tensor([[0.027417]], device=dev), "symmetric", True)
prec = mgp.PrecisionMaternOperator(lap, 2, torch.tensor([[0.5]], device=dev))
B = torch.randn(n, 16, device=dev, generator=torch.Generator(device=dev).manual_seed(1))
for _ in range(2):
    solvers.linear_cg(prec, B, tolerance=1e-6, max_iter=4000)
torch.cuda.synchronize()
for rep in range(3):
    t0 = time.perf_counter()
    pr = cProfile.Profile(); pr.enable()
    sol, info = solvers.linear_cg(prec, B, tolerance=1e-6, max_iter=4000, return_info=True)
    torch.cuda.synchronize()
    pr.disable()
    print("rep", rep, "wall_ms", round((time.perf_counter() - t0) * 1e3, 1), "iters", info["iterations"])
    if rep == 2:
        pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
