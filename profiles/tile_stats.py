#!/usr/bin/env python
"""Tile statistics of the v2 SpMM structure for the cfg-C graph (halo sizes, shared-memory footprint)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import manifold_gp_b200 as mgp
from manifold_gp_b200 import graph
from manifold_gp_b200.utils import synthetic

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
dev = torch.device("cuda:0")
x = synthetic.torus(n, seed=0, device=dev)
idx, val = mgp.NearestNeighbors(x).graph(32)
for name, perm in (("morton", getattr(idx, graph._PERM_ATTR)),):
    st = graph.GraphStructure(idx, n, perm=perm)
    t = st.build_tiles()
    hl = (t["halo_ptr"][1:] - t["halo_ptr"][:-1]).float()
    R = t["rows"]
    tstart = st.rowptr[torch.arange(0, n, R, device=dev)].long()
    tend = torch.cat([tstart[1:], st.rowptr[-1:].long()])
    nz = (tend - tstart).float()
    q = torch.tensor([0.5, 0.9, 0.99, 0.999, 1.0], device=dev)
    print(json.dumps({"order": name, "n": n, "tiles": int(hl.numel()), "halo_mean": float(hl.mean()),
                      "halo_quantiles_50_90_99_999_max": [float(v) for v in torch.quantile(hl, q)],
                      "nz_mean": float(nz.mean()), "nz_quantiles": [float(v) for v in torch.quantile(nz, q)],
                      "lmax": t["lmax"], "nzmax": t["nzmax"]}))
