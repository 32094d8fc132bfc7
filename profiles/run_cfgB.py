#!/usr/bin/env python
"""BASELINE cfg-B end to end on one B200: RMNIST-shape cloud (70k x 784, k = 10) -> tcgen05 kNN -> symmetrised graph ->
Laplacian -> Lanczos top-500 eigenpairs (RiemannMaternKernel.eval) -> spectral features / out-of-sample extension ->
semi-supervised posterior, with stage timings and self-consistency checks.  Development / evidence tool (bench.py is the
judged benchmark).   python profiles/run_cfgB.py [n] [d] [modes]"""
import json, os, sys, time, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import manifold_gp_b200 as mgp
from manifold_gp_b200 import _lib, solvers
from manifold_gp_b200._compat import gp as gpc
from manifold_gp_b200.utils import synthetic

warnings.simplefilter("ignore")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 70000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 784
modes = int(sys.argv[3]) if len(sys.argv) > 3 else 500
k, nu, n_test = 10, 2, 2000
dev = torch.device("cuda:0")
res = {"workload": f"rmnist_shape_N{n}_d{d}_k{k}_modes{modes}"}


def stage(name, fn):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize(); res[name + "_s"] = round(time.perf_counter() - t0, 4)
    return out


xall = synthetic.rmnist_shape(n + n_test, d, device=dev)
perm = torch.randperm(n + n_test, device=dev, generator=torch.Generator(device=dev).manual_seed(3))
x, xt = xall[perm[:n]].contiguous(), xall[perm[n:]].contiguous()
target = lambda z: torch.sin(3.0 * z[:, :8].sum(1)) + 0.5 * z[:, 8:16].sum(1)     # smooth on each 1-D family
y, yt = target(x), target(xt)
_lib.reset_launch_count()
kernel = stage("knn_graph", lambda: mgp.RiemannMaternKernel(nu=nu, x=x, nearest_neighbors=k, laplacian_normalization="symmetric",
                                                            num_modes=modes, bump_scale=10.0, bump_decay=1.0))
res["knn_kernel"] = kernel.knn.last_search["kernel"]
if "stats" in kernel.knn.last_search:
    st = kernel.knn.last_search["stats"].cpu()
    res["knn_research_queries"] = int(st[0])
res["edges_M"] = int(kernel.edge_index.shape[1])
kernel = kernel.to(dev)
d2, _ = kernel.knn.search(x[:4096].contiguous(), k)
eps = float(d2[:, k - 1].sqrt().median())
kernel.graphbandwidth = torch.tensor([[eps]], device=dev)
kernel.lengthscale = torch.tensor([[1.0]], device=dev)
res["eps"] = eps
lab = torch.zeros(n, dtype=torch.bool, device=dev)
lab[torch.randperm(n, device=dev, generator=torch.Generator(device=dev).manual_seed(4))[:n // 100]] = True
likelihood = gpc.GaussianLikelihood().to(dev)
likelihood.noise = torch.tensor([1e-2], device=dev)
model = mgp.RiemannGP(x[lab], y[lab], likelihood, gpc.ScaleKernel(kernel).to(dev), labeled=lab).to(dev)

# Lanczos top-`modes` eigenpairs (3 * modes steps, full re-orthogonalisation)
stage("lanczos_eigenpairs", lambda: model.eval())
lap = kernel.laplacian_operator._symmetric_twin()
phi = kernel.eigvec * kernel.laplacian_operator.degree_mat.pow(0.5).view(-1, 1)      # back to the symmetric operator's vectors
phi = phi / phi.norm(dim=0, keepdim=True)
sel = torch.tensor([1, 2, 5, 10, 50, min(100, modes - 1), min(250, modes - 1)], device=dev).unique()
r = lap._matmul(phi[:, sel].contiguous()) - phi[:, sel] * kernel.eigval[sel]
res["eig_residual_max"] = float(r.norm(dim=0).max())
res["eigval_first"] = [round(float(v), 6) for v in kernel.eigval[:6]]
res["eigval_last"] = float(kernel.eigval[-1])

# features on the graph and the out-of-sample extension at held-out points (kNN queries != database)
z = stage("features_graph", lambda: kernel.features(x))
zt = stage("features_out_of_sample", lambda: kernel.features(xt))
res["oos_knn_kernel"] = kernel.knn.last_search["kernel"]
res["oos_nonzero_rows"] = int((zt.abs().sum(1) > 0).sum())

# one Matern precision CG solve on this graph (16 right-hand sides)
prec = kernel.precision()
B = torch.randn(n, 16, device=dev, generator=torch.Generator(device=dev).manual_seed(1))
with torch.no_grad():
    sol, info = stage("precision_cg_16rhs", lambda: solvers.linear_cg(prec, B, tolerance=1e-6, max_iter=4000, return_info=True))
    res["cg_iterations"] = int(info["iterations"])
    res["cg_true_rel_residual"] = float(((prec._matmul(sol) - B).norm(dim=0) / B.norm(dim=0)).mean())

# semi-supervised posterior at the held-out points
try:
    with torch.no_grad():
        def post():
            model.posterior(xt)
            return model.posterior_mean
        mean = stage("posterior_mean", post)
        res["posterior_rmse"] = float((mean - yt).square().mean().sqrt())
        res["target_std"] = float(yt.std())
except Exception as e:   # report, do not hide
    res["posterior_error"] = repr(e)[:300]
res["gpu_launches"] = _lib.launch_count()
print(json.dumps(res))
