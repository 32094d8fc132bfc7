#!/usr/bin/env python
"""the semi-supervised Schur complement at scale.  One outer matvec of SchurComplementOperator
(Q_xx - Q_xz Q_zz^-1 Q_zx, schur_complement_operator.py:26-30) = 2 full-size precision matvecs + ONE inner CG solve on the
unlabelled block.  Times the round-2 fused inner solve (full index space, fused CUDA CG with CUDA graph, PrincipalBlockOperator)
against the generic path (MaskedOperator + un-fused CG through `_matmul`), reports inner iterations per outer matvec.
    python profiles/run_schur.py [n] [labelled fraction]"""
import json, os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import manifold_gp_b200 as mgp
from manifold_gp_b200 import solvers
from manifold_gp_b200.utils import synthetic

warnings.simplefilter("ignore")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.01
dev = torch.device("cuda:0")
x = synthetic.torus(n, seed=0, device=dev)
knn = mgp.NearestNeighbors(x)
idx, val = knn.graph(32)
d2, _ = knn.search(x[:4096].contiguous(), 32)
eps = float(d2[:, 31].sqrt().median())
lap = mgp.GraphLaplacianOperator(val, idx, n, torch.tensor([[eps]], device=dev), "symmetric", True)
prec = mgp.PrecisionMaternOperator(lap, 2, torch.tensor([[0.5]], device=dev))
mask = torch.zeros(n, dtype=torch.bool, device=dev)
mask[torch.randperm(n, device=dev, generator=torch.Generator(device=dev).manual_seed(4))[:int(n * frac)]] = True
V = torch.randn(int(mask.sum()), 16, device=dev, generator=torch.Generator(device=dev).manual_seed(1))
res = {"n": n, "labelled": int(mask.sum()), "rhs": 16, "cg_tolerance": 1e-4}
iters = []
orig = solvers.linear_cg
def spy(*a, **k):
    k2 = dict(k); k2["return_info"] = True
    out = orig(*a, **k2)
    iters.append(out[-1]["iterations"])
    return out if k.get("return_info") else out[0] if len(out) == 2 else out[:-1]
solvers.linear_cg = spy
outs = {}
for name, flag in (("fused_full_space_inner_solve", "1"), ("generic_masked_inner_solve", "0")):
    os.environ["MGP_FUSED_WRAPPERS"] = flag
    sch = mgp.SchurComplementOperator(prec, mask)
    with torch.no_grad(), mgp.settings.max_cholesky_size(0), mgp.settings.cg_tolerance(1e-4), mgp.settings.max_cg_iterations(4000):
        sch._matmul(V)                                   # warm-up: structure, layouts, graph capture
        torch.cuda.synchronize(); iters.clear(); t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            out = sch._matmul(V)
        torch.cuda.synchronize()
        res[name] = {"ms_per_outer_matvec": round((time.perf_counter() - t0) * 1e3 / reps, 2), "inner_cg_iterations": iters[-1] if iters else None}
        outs[name] = out
res["rel_diff_between_paths"] = float((outs["fused_full_space_inner_solve"] - outs["generic_masked_inner_solve"]).norm() / outs["generic_masked_inner_solve"].norm())
print(json.dumps(res))
