"""Device-side graph structure and thin Python wrappers over the C ABI (libmgp_b200.so).

``GraphStructure`` is the hyper-parameter independent part of the Laplacian: the row-major directed structure
(CSR over both directions of every edge of the reference's upper-triangular COO, nearest_neighbors.py:48-51) built
once per graph and cached on the ``idx`` tensor.  Per-bandwidth values are produced by :func:`lap_values`.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import c_int32, c_int64, c_size_t, ptr, stream


def _device_scalar(v, dtype, device) -> torch.Tensor:
    """0-d/1-element device tensor of ``dtype`` holding ``v`` (tensor -> no host sync; python number -> H2D of 1 value)."""
    if torch.is_tensor(v):
        return v.detach().reshape(-1)[:1].to(device=device, dtype=dtype).contiguous()
    return torch.tensor([float(v)], dtype=dtype, device=device)


class GraphStructure:
    """rowptr[n+1], col[nnz], eid[nnz] (int32) with nnz = 2M; column indices ascending inside a row."""

    def __init__(self, idx: torch.Tensor, n: int):
        if not idx.is_cuda:
            raise RuntimeError("GraphStructure: edge index must be a CUDA tensor (no CPU fallback exists)")
        if idx.dim() != 2 or idx.shape[0] != 2:
            raise ValueError("edge index must be [2, M]")
        idx = idx.to(torch.int64)
        if idx.stride(1) != 1:
            idx = idx.contiguous()
        self.n = int(n)
        self.m = int(idx.shape[1])
        self.nnz = 2 * self.m
        dev = idx.device
        self.device = dev
        self.rowptr = torch.empty(self.n + 1, dtype=torch.int32, device=dev)
        self.col = torch.empty(self.nnz, dtype=torch.int32, device=dev)
        self.eid = torch.empty(self.nnz, dtype=torch.int32, device=dev)
        ws_bytes = _lib.query("mgp_csr_build_ws_bytes", c_int64(self.n), c_int64(self.m))
        ws = _lib.workspace(ws_bytes, dev)
        _lib.call("mgp_csr_build", ptr(idx), c_int64(idx.stride(0)), c_int64(self.m), c_int64(self.n),
                  ptr(self.rowptr), ptr(self.col), ptr(self.eid), ptr(ws), c_size_t(ws.numel()), stream())
        self._d2 = {}
        self._dot_ws = None
        self._upper_pos = None

    # -- cached helpers -------------------------------------------------------------------------------------------
    def d2csr(self, val: torch.Tensor) -> torch.Tensor:
        """Per-directed-entry copy of the per-edge squared distances (cached per (storage, dtype))."""
        key = (val.data_ptr(), val.dtype, val._version)
        hit = self._d2.get(key)
        if hit is None:
            v = val.detach().reshape(-1).contiguous()
            out = torch.empty(self.nnz, dtype=v.dtype, device=self.device)
            _lib.call("mgp_gather_edge_" + _lib.suffix(v.dtype), ptr(v), ptr(self.eid), c_int64(self.nnz), ptr(out), stream())
            self._d2 = {key: out}  # keep only the latest
            hit = out
        return hit

    def dot_ws(self) -> torch.Tensor:
        if self._dot_ws is None:
            nb = _lib.query("mgp_lap_spmm_dot_ws_bytes", c_int64(self.n), c_int32(32))
            self._dot_ws = torch.zeros(nb, dtype=torch.uint8, device=self.device)
        return self._dot_ws

    def upper_pos(self) -> torch.Tensor:
        """Position in the CSR arrays of the (row<col) copy of every undirected edge e -- maps per-entry arrays back
        to the reference's per-edge arrays (laplacian_triu etc.).  For a diagonal COO entry either copy is returned."""
        if self._upper_pos is None:
            rows = torch.repeat_interleave(torch.arange(self.n, device=self.device, dtype=torch.int64),
                                           (self.rowptr[1:] - self.rowptr[:-1]).to(torch.int64))
            upper = rows <= self.col.to(torch.int64)
            pos = torch.nonzero(upper).squeeze(1)
            out = torch.empty(self.m, dtype=torch.int64, device=self.device)
            out[self.eid[pos].to(torch.int64)] = pos
            self._upper_pos = out
        return self._upper_pos


_STRUCT_ATTR = "_mgp_b200_structure"


def structure_for(idx: torch.Tensor, n: int) -> GraphStructure:
    """Build (once) and cache the GraphStructure on the edge-index tensor object."""
    st = getattr(idx, _STRUCT_ATTR, None)
    if st is None or st.n != int(n) or st.m != int(idx.shape[1]) or st.device != idx.device:
        st = GraphStructure(idx, n)
        try:
            setattr(idx, _STRUCT_ATTR, st)
        except Exception:  # pragma: no cover - tensors normally accept attributes
            pass
    return st


# ---- value build ------------------------------------------------------------------------------------------------
def lap_values(st: GraphStructure, d2csr: torch.Tensor, eps, self_loops: bool):
    """(deg_unnorm[n], deg[n], diag[n], a[nnz]) for bandwidth ``eps`` -- graph_laplacian_operator.py:52-106."""
    dt = d2csr.dtype
    dev = st.device
    eps_t = _device_scalar(eps, dt, dev)
    deg_un = torch.empty(st.n, dtype=dt, device=dev)
    deg = torch.empty(st.n, dtype=dt, device=dev)
    diag = torch.empty(st.n, dtype=dt, device=dev)
    a = torch.empty(st.nnz, dtype=dt, device=dev)
    _lib.call("mgp_lap_values_" + _lib.suffix(dt), ptr(st.rowptr), ptr(st.col), ptr(d2csr), c_int64(st.n), ptr(eps_t),
              c_int32(1 if self_loops else 0), ptr(deg_un), ptr(deg), ptr(diag), ptr(a), stream())
    return deg_un, deg, diag, a


# ---- SpMM ---------------------------------------------------------------------------------------------------------
def lap_spmm(st: GraphStructure, a, diag, x, shift=None, pre=None, post=None, out=None, dot_with=None, dot_out=None):
    """Y = post .* ((diag + shift) .* (pre .* X) - A (pre .* X)).  ``x``: [n, C] CUDA tensor with unit column stride."""
    if x.dim() != 2 or x.shape[0] != st.n:
        raise ValueError(f"lap_spmm: expected rhs of shape [{st.n}, C], got {tuple(x.shape)}")
    if x.stride(1) != 1:
        x = x.contiguous()
    dt = x.dtype
    if a.dtype != dt or diag.dtype != dt:
        raise TypeError(f"lap_spmm: value dtype {a.dtype} does not match rhs dtype {dt}")
    c = int(x.shape[1])
    if out is None:
        out = torch.empty((st.n, c), dtype=dt, device=x.device)
    shift_t = None if shift is None else _device_scalar(shift, dt, x.device)
    ws = st.dot_ws() if dot_out is not None else None
    _lib.call("mgp_lap_spmm_" + _lib.suffix(dt), ptr(st.rowptr), ptr(st.col), ptr(a), ptr(diag), ptr(shift_t),
              ptr(pre), ptr(post), ptr(x), c_int64(x.stride(0)), ptr(out), c_int64(out.stride(0)), c_int64(st.n),
              c_int32(c), ptr(dot_with), ptr(dot_out), ptr(ws), stream())
    return out


def lap_sddmm(st: GraphStructure, gy, x, pre=None, post=None):
    """(g_a[nnz], g_diag[n]) -- gradient of sum(gy * Y) w.r.t. the matrix entries of the SpMM above."""
    if gy.stride(1) != 1:
        gy = gy.contiguous()
    if x.stride(1) != 1:
        x = x.contiguous()
    dt = x.dtype
    g_a = torch.empty(st.nnz, dtype=dt, device=x.device)
    g_diag = torch.empty(st.n, dtype=dt, device=x.device)
    _lib.call("mgp_lap_sddmm_" + _lib.suffix(dt), ptr(st.rowptr), ptr(st.col), ptr(pre), ptr(post), ptr(gy),
              c_int64(gy.stride(0)), ptr(x), c_int64(x.stride(0)), c_int64(st.n), c_int32(int(x.shape[1])), ptr(g_a),
              ptr(g_diag), stream())
    return g_a, g_diag
