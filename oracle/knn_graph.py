"""Oracle: exact kNN search and the symmetrised, mean-coalesced kNN graph.  TEST INFRASTRUCTURE ONLY.

Restates ``manifold_gp/utils/nearest_neighbors.py`` (reference).  The arithmetic of the search and of
``coalesce`` lives in faiss / torch_sparse (third party, un-pinned ``setup.py:26-32``, absent here), so
those two are restated from their published behaviour -- PARITY UNPINNED against
the reference; pinned against fp64 brute force in ``tests/test_oracle_knn.py``.
"""

from __future__ import annotations

import torch


def knn_search(db: torch.Tensor, q: torch.Tensor, k: int, block: int = 2048, form: str = "direct"):
    """Exact k nearest neighbours of each row of ``q`` in ``db``: squared L2, ascending.

    Follows ``NearestNeighbors.search`` (nearest_neighbors.py:35-37) = ``faiss.Index*.search(x, k)``:
    returns ``(dist2[Q,k], idx[Q,k] int64)``.  The reference always scans exhaustively (FlatL2 for
    n < 5000, IVFFlat with ``nlist=1`` otherwise -- nearest_neighbors.py:12,23,25; riemann_kernel.py:40).

    ``form="direct"``  : ``sum_d (x_d - y_d)^2`` in the input dtype -- what faiss' IVF list scanner and the
                         small-batch Flat path compute (``fvec_L2sqr``).
    ``form="blas"``    : ``|x|^2 + |y|^2 - 2 x.y`` clamped at 0 -- faiss' large-batch Flat / GPU path.
    Ties are broken by ascending index (faiss' heap order on exact ties is unspecified).
    """
    assert db.dim() == 2 and q.dim() == 2 and db.shape[1] == q.shape[1]
    n = db.shape[0]
    k_eff = min(k, n)
    out_d = torch.empty(q.shape[0], k, dtype=db.dtype)
    out_i = torch.full((q.shape[0], k), -1, dtype=torch.int64)
    db_sq = (db * db).sum(1)
    for s in range(0, q.shape[0], block):
        qb = q[s:s + block]
        if form == "blas":
            d2 = (qb * qb).sum(1, keepdim=True) + db_sq.unsqueeze(0) - 2.0 * (qb @ db.T)
            d2.clamp_(min=0)
        else:
            d2 = torch.zeros(qb.shape[0], n, dtype=db.dtype)
            for j in range(db.shape[1]):  # sequential over d: the summation order the CUDA kernel uses for small d
                diff = qb[:, j:j + 1] - db[:, j].unsqueeze(0)
                sq = diff * diff          # rounded product, then a rounded add: no FMA contraction (the CUDA
                d2 += sq                  # kernel uses __fmul_rn / __fadd_rn so the two agree bit for bit)
        # stable sort on (distance, index): ascending distance, ties by ascending index
        dd, ii = torch.sort(d2, dim=1, stable=True)
        out_d[s:s + block, :k_eff] = dd[:, :k_eff]
        out_i[s:s + block, :k_eff] = ii[:, :k_eff]
    if k_eff < k:
        out_d[:, k_eff:] = float("inf")
    return out_d, out_i


def knn_search_exact(db: torch.Tensor, q: torch.Tensor, k: int, block: int = 1024):
    """fp64 brute-force ground truth used to pin ``knn_search`` and the CUDA kernel (set equality up to ties)."""
    d, i = knn_search(db.double(), q.double(), k, block=block, form="direct")
    return d, i


def symmetrize_coalesce(dist2: torch.Tensor, idx: torch.Tensor, n: int, drop_first: bool = True):
    """Directed kNN lists -> upper-triangular, lexicographically sorted, mean-coalesced COO.

    Follows ``NearestNeighbors.graph`` (nearest_neighbors.py:39-55):
      * ``:42-43``  drop column 0 (the query itself) unless ``self_loop``;
      * ``:45-46``  rows = arange(N).repeat_interleave(k-1), cols = idx.flatten();
      * ``:48-50``  ``split = cols > rows``; edge (r,c) -> (min,max)   [a surviving ``c == r`` stays (r,r)];
      * ``:51``     ``torch_sparse.coalesce(..., op='mean')``: sort by ``row*n+col``, average duplicates.
    Returns ``(idx[2,M] int64, val[M])``.
    """
    if drop_first:
        dist2, idx = dist2[:, 1:], idx[:, 1:]
    kk = idx.shape[1]
    rows = torch.arange(idx.shape[0], dtype=torch.int64).repeat_interleave(kk)
    cols = idx.reshape(-1).to(torch.int64)
    val = dist2.reshape(-1)
    split = cols > rows
    r = torch.where(split, rows, cols)
    c = torch.where(split, cols, rows)
    key = r * n + c
    ukey, inv, cnt = torch.unique(key, return_inverse=True, return_counts=True)
    acc = torch.zeros(ukey.shape[0], dtype=val.dtype).scatter_add_(0, inv, val)
    out_val = acc / cnt.to(val.dtype)
    out_idx = torch.stack([ukey // n, ukey % n], dim=0)
    return out_idx, out_val


def knn_graph(x: torch.Tensor, k: int, form: str = "direct", block: int = 2048):
    """``NearestNeighbors(x).graph(k)`` (nearest_neighbors.py:39-55): k-1 out-edges per node, symmetrised."""
    d2, idx = knn_search(x, x, k, block=block, form=form)
    return symmetrize_coalesce(d2, idx, x.shape[0], drop_first=True)
