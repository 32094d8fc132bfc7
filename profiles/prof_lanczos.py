"""Where does a Lanczos run spend its time?  cProfile of solvers.lanczos_tridiag on an RMNIST-shape graph."""
import cProfile, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import manifold_gp_b200 as mgp
from manifold_gp_b200 import solvers, _lib
from manifold_gp_b200.utils import synthetic

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 300
dev = torch.device("cuda:0")
x = synthetic.rmnist_shape(n, 784, device=dev)
idx, val = mgp.NearestNeighbors(x).graph(10)
lap = mgp.GraphLaplacianOperator(val, idx, n, torch.tensor([[0.19]], device=dev), "symmetric", True)
solvers.lanczos_tridiag(lap, 10)
torch.cuda.synchronize()
_lib.reset_launch_count()
t0 = time.perf_counter()
pr = cProfile.Profile()
pr.enable()
q, t = solvers.lanczos_tridiag(lap, steps)
torch.cuda.synchronize()
pr.disable()
print("steps", q.shape[0], "seconds", round(time.perf_counter() - t0, 3), "launches", _lib.launch_count())
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
