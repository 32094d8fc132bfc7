"""``NearestNeighbors`` -- drop-in for manifold_gp/utils/nearest_neighbors.py on B200.

Same surface (``train / search / graph / min_ivf``, ``.x``); faiss and torch_sparse.coalesce are replaced by
``mgp_knn_search`` (exact brute force -- the reference's IVF index has ``nlist=1`` and is therefore exhaustive too,
nearest_neighbors.py:12,23,25 + riemann_kernel.py:40) and ``mgp_graph_symmetrize`` (radix sort + segmented mean).
"""
from __future__ import annotations

import os

import torch

from .. import _lib
from .._lib import c_int32, c_int64, c_size_t, ptr, stream


class NearestNeighbors():
    def __init__(self, x=None, nlist=1) -> None:
        self.min_ivf = 5000
        # use the tcgen05 search (results are identical either way; MGP_KNN_TC=0 selects the CUDA-core kernel for A/B runs)
        self.tensor_core = os.environ.get("MGP_KNN_TC", "1") != "0"
        self.last_search = None
        if x is not None:
            self.train(x, nlist)

    def train(self, x, nlist=1):
        """The reference builds a faiss index here (:17-33); the brute-force kernel needs only the points."""
        if not x.is_cuda:
            raise RuntimeError("NearestNeighbors: x must be a CUDA tensor (manifold_gp_b200 has no CPU fallback)")
        self.x = x
        self._db = x.detach().to(torch.float32).contiguous()
        self.nlist = nlist
        return self

    def search(self, x, k, nprobe=1):
        """(dist2[Q,k] ascending squared L2, idx[Q,k] int64) -- :35-37."""
        if not x.is_cuda:
            raise RuntimeError("NearestNeighbors.search: x must be a CUDA tensor (no CPU fallback exists)")
        q = x.detach().to(torch.float32).contiguous()
        n, d = self._db.shape
        nq = q.shape[0]
        if q.shape[1] != d:
            raise ValueError(f"query dimension {q.shape[1]} != database dimension {d}")
        dist = torch.empty((nq, k), dtype=torch.float32, device=q.device)
        idx = torch.empty((nq, k), dtype=torch.int64, device=q.device)
        same = q.data_ptr() == self._db.data_ptr() and nq == n
        # multi-GPU: with a distributed backend registered the self-search of a large cloud is sharded by query
        # rows (database replicated, no collective in the search) and the lists are all-gathered
        from .. import solvers as _solvers
        be = _solvers._DIST_BACKEND
        if same and be is not None and getattr(be, "world", 1) > 1 and n >= be.min_rows and not getattr(self, "_in_shard", False):
            self._in_shard = True
            try:
                return be.sharded_self_search(self, k)
            finally:
                self._in_shard = False
        self.last_search = {"kernel": "cuda_core"}
        if self.tensor_core:
            # k <= 48, n >= 256: TF32 tcgen05 distance tiles + fused top-(k+margin) + exact certified re-rank (knn_tc.cu);
            # bit-identical results to the CUDA-core kernel below.
            wide = self._wide_window(q, k) if (nq >= 65536 and not getattr(self, "_in_pilot", False)) else False
            _lib.call("mgp_knn_tc_config", c_int32(64 if wide else 0))
            nb = _lib.query("mgp_knn_search_tc_ws_bytes", c_int64(n), c_int64(nq), c_int32(d), c_int32(k), c_int32(int(same)))
            if nb <= 0:
                _lib.call("mgp_knn_tc_config", c_int32(0))
            if nb > 0:
                ws = _lib.workspace(nb, q.device)
                stats = torch.zeros(4, dtype=torch.int32, device=q.device)
                rc = _lib.call_rc("mgp_knn_search_tc_f32", ptr(self._db), c_int64(n), ptr(self._db if same else q), c_int64(nq),
                                  c_int32(d), c_int32(k), ptr(dist), ptr(idx), ptr(ws), c_size_t(ws.numel()), ptr(stats), stream())
                _lib.call("mgp_knn_tc_config", c_int32(0))
                if rc == 0:
                    self.last_search = {"kernel": "tcgen05", "stats": stats, "wide_window": bool(wide)}   # stats: device tensor
                    return dist.to(self.x.dtype), idx
                if rc != _lib.MGP_EUNSUPPORTED:
                    raise RuntimeError(f"mgp_knn_search_tc_f32 failed ({rc}): {_lib.last_error()}")
        nb = _lib.query("mgp_knn_search_ws_bytes", c_int64(n), c_int64(nq), c_int32(d), c_int32(k))
        ws = _lib.workspace(nb, q.device)
        _lib.call("mgp_knn_search_f32", ptr(self._db), c_int64(n), ptr(q), c_int64(nq), c_int32(d), c_int32(k),
                  ptr(dist), ptr(idx), ptr(ws), c_size_t(ws.numel()), stream())
        return dist.to(self.x.dtype), idx

    def _wide_window(self, q, k) -> bool:
        """Pilot for large searches: 1024 evenly spaced queries through the tensor-core search with the default re-rank window.
        If more than 2 % of them fail the exactness certificate (and would be re-searched exhaustively on the CUDA cores), the
        full search keeps a 64-candidate window instead.  Results are exact either way; this only chooses the faster route.
        Costs one small search (~10 ms at N = 1M, d = 3) and one 16-byte device read."""
        nq = q.shape[0]
        if k > 16:                   # k + 16 already rounds to a 64-candidate window
            return False
        sel = q[:: max(1, nq // 1024)][:1024].contiguous()
        self._in_pilot = True
        try:
            self.search(sel, k)
            info = self.last_search
        finally:
            self._in_pilot = False
        if info.get("kernel") != "tcgen05":
            return False
        failed = int(info["stats"][0])
        return failed > 0.02 * sel.shape[0]

    def graph(self, k, symmetric=True, self_loop=False, nprobe=1):
        """(idx[2,M] int64 upper-triangular sorted, val[M] mean squared distance) -- :39-55."""
        val, idx = self.search(self.x, k, nprobe)
        n = self.x.shape[0]
        if not symmetric:
            if not self_loop:
                val, idx = val[:, 1:], idx[:, 1:]
            rows = torch.arange(n, device=idx.device).repeat_interleave(idx.shape[1])
            return torch.stack([rows, idx.reshape(-1)], dim=0), val.reshape(-1)
        drop = 0 if self_loop else 1
        cap = n * (k - drop)
        dev = idx.device
        eidx = torch.empty((2, cap), dtype=torch.int64, device=dev)
        ev = torch.empty(cap, dtype=torch.float32, device=dev)
        m_out = torch.zeros(1, dtype=torch.int64, device=dev)
        nb = _lib.query("mgp_graph_symmetrize_ws_bytes", c_int64(n), c_int32(k))
        ws = _lib.workspace(nb, dev)
        v32 = val.to(torch.float32).contiguous()
        _lib.call("mgp_graph_symmetrize_f32", ptr(v32), ptr(idx), c_int64(n), c_int32(k), c_int32(drop), ptr(eidx),
                  ptr(ev), c_int64(cap), ptr(m_out), ptr(ws), c_size_t(ws.numel()), stream())
        m = int(m_out.item())          # one host read per graph build (output size is data dependent)
        out_idx = eidx[:, :m].contiguous()
        # hint for the operators: a space-filling-curve order of the points (the reference's outputs are unchanged)
        from .. import graph
        graph.attach_permutation(out_idx, graph.morton_permutation(self.x))
        return out_idx, ev[:m].to(self.x.dtype).contiguous()

    @property
    def min_ivf(self):
        return self._min_ivf

    @min_ivf.setter
    def min_ivf(self, value):
        self._min_ivf = value
