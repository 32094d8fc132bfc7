"""CPU statistics of the quad-row stream layout (graph.quad_union_lists / quad_streams) on a 200k-point torus graph, k = 32:
union columns, padding, HBM bytes and shared-memory wavefronts per nonzero.  python profiles/quad_stream_stats.py"""
import math, sys, numpy as np, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from sklearn.neighbors import NearestNeighbors
from manifold_gp_b200 import graph
n, k, R = 200_000, 32, 128
rng = np.random.default_rng(0)
u, v = rng.random(n) * 2 * math.pi, rng.random(n) * 2 * math.pi
x = np.stack(((2 + np.cos(v)) * np.cos(u), (2 + np.cos(v)) * np.sin(u), np.sin(v)), 1).astype(np.float32)
_, idx = NearestNeighbors(n_neighbors=k).fit(x).kneighbors(x)
rows = np.repeat(np.arange(n), k - 1); cols = idx[:, 1:].reshape(-1)
key = np.unique(np.concatenate([rows, cols]).astype(np.int64) * n + np.concatenate([cols, rows]))
a, b = torch.from_numpy(key // n), torch.from_numpy(key % n)
perm = graph.morton_permutation(torch.from_numpy(x))
inv = torch.empty_like(perm); inv[perm] = torch.arange(n)
r, c = inv[a], inv[b]
o = torch.argsort(r * n + c); r, c = r[o], c[o]
rowptr = torch.zeros(n + 1, dtype=torch.int64); rowptr[1:] = torch.cumsum(torch.bincount(r, minlength=n), 0)
q = graph.quad_union_lists(rowptr, c, n, R)
lcol = torch.zeros_like(c)   # indices irrelevant for the statistics
s = graph.quad_streams(q, lcol, n, R, 16)
nnz = r.numel()
print({"nnz": nnz, "union_per_nonzero": round(q["union_per_nonzero"], 4), "stream_entries_per_nonzero": round(s["entries"] / nnz, 4),
       "padding": round(s["padding"], 4), "hbm_bytes_per_nonzero_idx_plus_values": round(s["entries"] * 18 / nnz, 3),
       "smem_wavefronts_per_nonzero": round(s["entries"] / 8 * 5.5 / nnz, 4)})
