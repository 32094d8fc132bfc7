import json, sys, os
sys.path.insert(0, os.getcwd())
import torch
import manifold_gp_b200 as mgp
from manifold_gp_b200.utils import synthetic
dev = torch.device("cuda:0")
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
out = []
for n, d, k in ((1_000_000, 64, 8), (1_000_000, 64, 32), (1_000_000, 784, 8), (1_000_000, 784, 32), (1_000_000, 3, 32)):
    x = synthetic.torus(n, device=dev) if d == 3 else synthetic.rmnist_shape(n, d, device=dev)
    knn = mgp.NearestNeighbors(x)
    knn.search(x[:2048].contiguous(), k)
    torch.cuda.synchronize(); ev0.record(); knn.search(x, k); ev1.record(); torch.cuda.synchronize()
    info = knn.last_search
    t = ev0.elapsed_time(ev1) * 1e-3
    out.append({"n": n, "d": d, "k": k, "knn_search_s_incl_pilot": round(t, 4), "kernel": info["kernel"], "wide_window": info.get("wide_window"),
                "research_queries": int(info["stats"][0]) if "stats" in info else None, "useful_tflops": round(2.0 * n * n * d / t / 1e12, 2)})
    print(json.dumps(out[-1]), flush=True)
    del x, knn
    torch.cuda.empty_cache()
