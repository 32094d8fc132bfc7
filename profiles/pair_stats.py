#!/usr/bin/env python
"""CPU analysis for the next SpMM generation: how many X-row loads would a paired-row /
quad-row stream save?  For Morton-adjacent rows r1, r2 of the symmetrised kNN graph of a torus cloud: |N(r1) u N(r2)| against
|N(r1)| + |N(r2)|.  Pure numpy / scikit-learn (no GPU); the graph statistics are scale free, so N = 200k stands in for 1M.
    python profiles/pair_stats.py [n] [k]"""
import json, math, sys
import numpy as np
from sklearn.neighbors import NearestNeighbors

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 32
rng = np.random.default_rng(0)
u, v = rng.random(n) * 2 * math.pi, rng.random(n) * 2 * math.pi
x = np.stack(((2 + np.cos(v)) * np.cos(u), (2 + np.cos(v)) * np.sin(u), np.sin(v)), 1).astype(np.float32)
nn = NearestNeighbors(n_neighbors=k).fit(x)
_, idx = nn.kneighbors(x)
rows = np.repeat(np.arange(n), k - 1)
cols = idx[:, 1:].reshape(-1)
# symmetrise: union of both directions, no self loops
a = np.concatenate([rows, cols]); b = np.concatenate([cols, rows])
key = np.unique(a.astype(np.int64) * n + b)
a, b = key // n, key % n
# Morton order over the 3 coordinates (10 bits each)
q = ((x - x.min(0)) / (x.max(0) - x.min(0) + 1e-9) * 1023).astype(np.int64)
def spread(t):
    t = (t | (t << 16)) & 0x030000FF; t = (t | (t << 8)) & 0x0300F00F
    t = (t | (t << 4)) & 0x030C30C3; t = (t | (t << 2)) & 0x09249249
    return t
code = spread(q[:, 0]) | (spread(q[:, 1]) << 1) | (spread(q[:, 2]) << 2)
perm = np.argsort(code, kind="stable")
inv = np.empty(n, np.int64); inv[perm] = np.arange(n)
a, b = inv[a], inv[b]
order = np.lexsort((b, a)); a, b = a[order], b[order]
ptr = np.searchsorted(a, np.arange(n + 1))
nbr = [b[ptr[i]:ptr[i + 1]] for i in range(n)]
res = {"n": n, "k": k, "nnz": int(len(a)), "mean_row_nnz": len(a) / n}
for g in (2, 4, 8):
    tot, uni = 0, 0
    for s in range(0, n - g + 1, g):
        sets = nbr[s:s + g]
        tot += sum(len(t) for t in sets)
        uni += len(np.unique(np.concatenate(sets)))
    res[f"group{g}"] = {"x_row_loads_per_nonzero": round(uni / tot, 4), "value_slots_per_nonzero": round(uni * g / tot, 4),
                        "note": "x_row_loads: shared-memory X-row loads relative to the row-by-row walk; value_slots: "
                                "dense g-wide value stream per union column relative to nnz (HBM bytes if stored unpacked)"}
# tile-level halo statistics for 128-row tiles
T = 128
halo = []
for t0 in range(0, n, T):
    cols_t = np.unique(np.concatenate(nbr[t0:t0 + T]))
    halo.append(int(((cols_t < t0) | (cols_t >= t0 + T)).sum()))
res["halo_rows_per_tile_mean"] = float(np.mean(halo)); res["halo_rows_per_tile_max"] = int(np.max(halo))
print(json.dumps(res))
