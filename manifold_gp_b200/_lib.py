"""ctypes binding of libmgp_b200.so (the C ABI declared in include/mgp_b200.h).

There is no CPU fallback: if the shared library is missing this module raises at import, and every call checks
that its tensors are CUDA tensors.  PyTorch is used for device memory and streams only; all arithmetic on the hot
path happens in the hand-written sm_100a kernels behind these entry points.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int32, c_int64, c_size_t, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmgp_b200.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: build it with `python manifold_gp_b200/csrc/build.py` "
        "(or `python -c 'import __graft_entry__ as g; g.build()'`).  manifold_gp_b200 has no CPU / PyTorch fallback."
    )

_dll = ctypes.CDLL(LIB_PATH)

P = c_void_p
_SIG = {
    "mgp_last_error": (c_char_p, []),
    "mgp_version": (c_char_p, []),
    "mgp_launch_count": (c_int64, []),
    "mgp_reset_launch_count": (None, []),
    "mgp_add_launch_count": (None, [c_int64]),
    "mgp_knn_search_ws_bytes": (c_size_t, [c_int64, c_int64, c_int32, c_int32]),
    "mgp_knn_search_f32": (c_int32, [P, c_int64, P, c_int64, c_int32, c_int32, P, P, P, c_size_t, P]),
    "mgp_knn_tc_config": (c_int32, [c_int32]),
    "mgp_knn_search_tc_ws_bytes": (c_size_t, [c_int64, c_int64, c_int32, c_int32, c_int32]),
    "mgp_knn_search_tc_f32": (c_int32, [P, c_int64, P, c_int64, c_int32, c_int32, P, P, P, c_size_t, P, P]),
    "mgp_graph_symmetrize_ws_bytes": (c_size_t, [c_int64, c_int32]),
    "mgp_graph_symmetrize_f32": (c_int32, [P, P, c_int64, c_int32, c_int32, P, P, c_int64, P, P, c_size_t, P]),
    "mgp_csr_build_ws_bytes": (c_size_t, [c_int64, c_int64]),
    "mgp_csr_build": (c_int32, [P, c_int64, c_int64, c_int64, P, P, P, P, c_size_t, P]),
    "mgp_gather_edge_f32": (c_int32, [P, P, c_int64, P, P]),
    "mgp_gather_edge_f64": (c_int32, [P, P, c_int64, P, P]),
    "mgp_lap_values_f32": (c_int32, [P, P, P, c_int64, P, c_int32, P, P, P, P, P]),
    "mgp_lap_values_f64": (c_int32, [P, P, P, c_int64, P, c_int32, P, P, P, P, P]),
    "mgp_lap_values_grad_ws_bytes": (c_size_t, [c_int64]),
    "mgp_lap_values_grad_f32": (c_int32, [P, P, P, c_int64, P, c_int32, P, P, P, P, P, P, P, P, P, P, P]),
    "mgp_lap_values_grad_f64": (c_int32, [P, P, P, c_int64, P, c_int32, P, P, P, P, P, P, P, P, P, P, P]),
    "mgp_lap_spmm_dot_ws_bytes": (c_size_t, [c_int64, c_int32]),
    "mgp_lap_spmm_f32": (c_int32, [P, P, P, P, P, P, P, P, c_int64, P, c_int64, c_int64, c_int32, P, P, P, P]),
    "mgp_lap_spmm_f64": (c_int32, [P, P, P, P, P, P, P, P, c_int64, P, c_int64, c_int64, c_int32, P, P, P, P]),
    "mgp_lap_spmm_tiled_f32": (c_int32, [P, P, P, P, P, P, c_int32, c_int32, c_int32, P, P, P, P, P, P, c_int64, P, c_int64, c_int64, c_int32, P, P, P, P]),
    "mgp_lap_spmm_tiled_f64": (c_int32, [P, P, P, P, P, P, c_int32, c_int32, c_int32, P, P, P, P, P, P, c_int64, P, c_int64, c_int64, c_int32, P, P, P, P]),
    "mgp_lap_pad_values_f32": (c_int32, [P, P, P, c_int64, P, P]),
    "mgp_lap_pad_values_f64": (c_int32, [P, P, P, c_int64, P, P]),
    "mgp_lap_spmm_pipe_f32": (c_int32, [P, P, P, P, P, P, c_int32, c_int32, c_int32, P, P, P, P, P, c_int64, P, c_int64, c_int64, c_int32, P, P, P, P]),
    "mgp_lap_spmm_pipe_f64": (c_int32, [P, P, P, P, P, P, c_int32, c_int32, c_int32, P, P, P, P, P, c_int64, P, c_int64, c_int64, c_int32, P, P, P, P]),
    "mgp_lap_spmv_tile_f32": (c_int32, [P, P, P, P, P, P, c_int32, c_int32, P, P, P, P, P, c_int64, P, c_int64, c_int64, P, P, P, P]),
    "mgp_lap_spmv_tile_f64": (c_int32, [P, P, P, P, P, P, c_int32, c_int32, P, P, P, P, P, c_int64, P, c_int64, c_int64, P, P, P, P]),
    "mgp_lap_wi_values_f32": (c_int32, [P, P, P, c_int64, P, P]),
    "mgp_lap_wi_values_f64": (c_int32, [P, P, P, c_int64, P, P]),
    "mgp_lap_spmm_wi_f32": (c_int32, [P, P, P, P, P, P, c_int32, c_int32, c_int32, c_int32, P, P, P, P, P, c_int64, P, c_int64, c_int64, c_int32, P, P, P, P, c_int32, c_int32, P, P, P]),
    "mgp_lap_spmm_wi_f64": (c_int32, [P, P, P, P, P, P, c_int32, c_int32, c_int32, c_int32, P, P, P, P, P, c_int64, P, c_int64, c_int64, c_int32, P, P, P, P, c_int32, c_int32, P, P, P]),
    "mgp_lap_spmm_wi_ex_f32": (c_int32, [P, P, P, P, P, P, c_int32, c_int32, c_int32, c_int32, P, P, P, P, P, c_int64, P, c_int64, c_int64, c_int32, P, P, P, P, c_int32, c_int32, P, P, P]),
    "mgp_lap_spmm_wi_ex_f64": (c_int32, [P, P, P, P, P, P, c_int32, c_int32, c_int32, c_int32, P, P, P, P, P, c_int64, P, c_int64, c_int64, c_int32, P, P, P, P, c_int32, c_int32, P, P, P]),
    "mgp_lap_sddmm_f32": (c_int32, [P, P, P, P, P, c_int64, P, c_int64, c_int64, c_int32, P, P, P]),
    "mgp_lap_sddmm_f64": (c_int32, [P, P, P, P, P, c_int64, P, c_int64, c_int64, c_int32, P, P, P]),
    "mgp_cg_state_elems": (c_size_t, [c_int32]),
    "mgp_cg_ws_bytes": (c_size_t, [c_int64, c_int32]),
    "mgp_cg_init_f32": (c_int32, [P, c_int64, P, P, P, c_int64, c_int64, c_int32, c_float, c_float, c_float, c_int32, c_int32, P, P, P]),
    "mgp_cg_init_f64": (c_int32, [P, c_int64, P, P, P, c_int64, c_int64, c_int32, c_double, c_double, c_double, c_int32, c_int32, P, P, P]),
    "mgp_cg_alpha_f32": (c_int32, [P, P, c_int64, c_int64, c_int32, c_int32, P, P, P]),
    "mgp_cg_alpha_f64": (c_int32, [P, P, c_int64, c_int64, c_int32, c_int32, P, P, P]),
    "mgp_cg_update_f32": (c_int32, [P, P, P, P, c_int64, c_int64, c_int32, P, P, c_int32, P, P]),
    "mgp_cg_update_f64": (c_int32, [P, P, P, P, c_int64, c_int64, c_int32, P, P, c_int32, P, P]),
    "mgp_cg_rupdate_f32": (c_int32, [P, P, c_int64, c_int64, c_int32, P, P, c_int32, P, P, P]),
    "mgp_cg_rupdate_f64": (c_int32, [P, P, c_int64, c_int64, c_int32, P, P, c_int32, P, P, P]),
    "mgp_cg_pxupdate_f32": (c_int32, [P, P, P, c_int64, c_int64, c_int32, P, P]),
    "mgp_cg_pxupdate_f64": (c_int32, [P, P, P, c_int64, c_int64, c_int32, P, P]),
    "mgp_cg_dist_norm2_f32": (c_int32, [P, c_int64, c_int64, c_int32, P, P, P, P]),
    "mgp_cg_dist_norm2_f64": (c_int32, [P, c_int64, c_int64, c_int32, P, P, P, P]),
    "mgp_cg_dist_init_f32": (c_int32, [P, c_int64, P, P, P, c_int64, c_int64, c_int32, P, P, P, P]),
    "mgp_cg_dist_init_f64": (c_int32, [P, c_int64, P, P, P, c_int64, c_int64, c_int32, P, P, P, P]),
    "mgp_cg_dist_update_f32": (c_int32, [P, P, P, P, c_int64, c_int64, c_int32, P, P, P, P]),
    "mgp_cg_dist_update_f64": (c_int32, [P, P, P, P, c_int64, c_int64, c_int32, P, P, P, P]),
    "mgp_cg_dist_scalars_f32": (c_int32, [P, P, c_int32, c_int32, c_float, c_float, c_float, c_int32, c_int32, P, c_int32, P]),
    "mgp_cg_dist_scalars_f64": (c_int32, [P, P, c_int32, c_int32, c_double, c_double, c_double, c_int32, c_int32, P, c_int32, P]),
    "mgp_cg_peer_scalars_f32": (c_int32, [P, P, c_int32, c_int32, c_float, c_float, c_float, c_int32, c_int32, P, c_int32, P, P, P, c_int32, c_int32, P]),
    "mgp_cg_peer_scalars_f64": (c_int32, [P, P, c_int32, c_int32, c_double, c_double, c_double, c_int32, c_int32, P, c_int32, P, P, P, c_int32, c_int32, P]),
    "mgp_cg_peer_rupdate_f32": (c_int32, [P, P, c_int64, c_int64, c_int32, P, P, P, P, P, P, c_int32, c_int32, P]),
    "mgp_cg_peer_rupdate_f64": (c_int32, [P, P, c_int64, c_int64, c_int32, P, P, P, P, P, P, c_int32, c_int32, P]),
    "mgp_cg_peer_pxupdate_f32": (c_int32, [P, P, P, c_int64, c_int64, c_int32, P, P, c_int32, P, P, P, c_int32, c_int32, P]),
    "mgp_cg_peer_pxupdate_f64": (c_int32, [P, P, P, c_int64, c_int64, c_int32, P, P, c_int32, P, P, P, c_int32, c_int32, P]),
    "mgp_peer_barrier": (c_int32, [P, P, c_int32, c_int32, P]),
    "mgp_peer_publish": (c_int32, [P, ctypes.c_uint32, c_int32, c_int32, P]),
    "mgp_cg_peer_cgstep_f32": (c_int32, [P, P, P, P, P, c_int64, c_int64, c_int32, P, P, c_int32, P, P, P, P, P, P, c_int32, c_int32, P]),
    "mgp_cg_peer_cgstep_f64": (c_int32, [P, P, P, P, P, c_int64, c_int64, c_int32, P, P, c_int32, P, P, P, P, P, P, c_int32, c_int32, P]),
    "mgp_cg_pupdate_f32": (c_int32, [P, P, c_int64, c_int64, c_int32, P, P]),
    "mgp_cg_pupdate_f64": (c_int32, [P, P, c_int64, c_int64, c_int32, P, P]),
    "mgp_cg_finalize_f32": (c_int32, [P, c_int64, P, c_int64, c_int64, c_int32, P, P]),
    "mgp_cg_finalize_f64": (c_int32, [P, c_int64, P, c_int64, c_int64, c_int32, P, P]),
    "mgp_lanczos_ws_bytes": (c_size_t, [c_int64, c_int32]),
    "mgp_lanczos_reorth_f32": (c_int32, [P, c_int64, c_int32, P, c_int64, P, P, P, P]),
    "mgp_lanczos_reorth_f64": (c_int32, [P, c_int64, c_int32, P, c_int64, P, P, P, P]),
    "mgp_lap_values_pass_f32": (c_int32, [c_int32, P, P, P, c_int64, P, c_int32, P, P, P, P, P]),
    "mgp_lap_values_pass_f64": (c_int32, [c_int32, P, P, P, c_int64, P, c_int32, P, P, P, P, P]),
    "mgp_lap_pair_values_f32": (c_int32, [P, P, c_int64, P, P]),
    "mgp_lap_pair_values_f64": (c_int32, [P, P, c_int64, P, P]),
    "mgp_lanczos_axpy_f32": (c_int32, [P, c_int64, c_int32, P, c_int64, P, P, P, P]),
    "mgp_lanczos_axpy_f64": (c_int32, [P, c_int64, c_int32, P, c_int64, P, P, P, P]),
    "mgp_lanczos_normalize_f32": (c_int32, [P, c_int64, P, P, P, P]),
    "mgp_lanczos_normalize_f64": (c_int32, [P, c_int64, P, P, P, P]),
    "mgp_lanczos_dots_f32": (c_int32, [P, c_int64, c_int32, P, c_int64, P, P, P]),
    "mgp_lanczos_dots_f64": (c_int32, [P, c_int64, c_int32, P, c_int64, P, P, P]),
    "mgp_out_of_sample_f32": (c_int32, [P, P, c_int64, c_int32, P, P, P, c_int32, P, c_int64, c_int32, P, c_int64, P]),
    "mgp_out_of_sample_f64": (c_int32, [P, P, c_int64, c_int32, P, P, P, c_int32, P, c_int64, c_int32, P, c_int64, P]),
}

class WiExt(ctypes.Structure):
    """``mgp_wi_ext`` of include/mgp_b200.h (optional hooks of mgp_lap_spmm_wi_ex)."""
    _fields_ = [("done_flag", c_void_p), ("wait_flags", c_void_p), ("publish_flags", c_void_p), ("ticket", c_void_p),
                ("red_ptrs", c_void_p), ("red_flags", c_void_p), ("ship_extra", c_void_p), ("ship_ncols", c_int32),
                ("ep_add", c_int32), ("ep_coef", c_void_p), ("publish_at_start", c_int32), ("reserved", c_int32),
                ("pair_rows", c_void_p)]


def wi_ext(done_flag=None, wait_flags=None, publish_flags=None, ticket=None, red_ptrs=None, red_flags=None, ship_extra=None,
           ship_ncols=0, ep_coef=None, ep_add=False, publish_at_start=False, pair_rows=None):
    def dp(t):
        if t is None:
            return None
        if not t.is_cuda:
            raise RuntimeError("manifold_gp_b200: expected a CUDA tensor (no CPU fallback exists)")
        return t.data_ptr()
    return WiExt(dp(done_flag), dp(wait_flags), dp(publish_flags), dp(ticket), dp(red_ptrs), dp(red_flags), dp(ship_extra),
                 int(ship_ncols), 1 if ep_add else 0, dp(ep_coef), 1 if publish_at_start else 0, 0, dp(pair_rows))


for _name, (_res, _args) in _SIG.items():
    _fn = getattr(_dll, _name)  # AttributeError here = header / library mismatch: fail loudly
    _fn.restype = _res
    _fn.argtypes = _args

EXPORTS = tuple(_SIG)


def version() -> str:
    return _dll.mgp_version().decode()


def launch_count() -> int:
    return int(_dll.mgp_launch_count())


def reset_launch_count() -> None:
    _dll.mgp_reset_launch_count()


def last_error() -> str:
    return _dll.mgp_last_error().decode()


def suffix(dtype: torch.dtype) -> str:
    if dtype == torch.float32:
        return "f32"
    if dtype == torch.float64:
        return "f64"
    raise TypeError(f"manifold_gp_b200 kernels support float32 and float64, got {dtype}")


def ptr(t):
    """Device pointer of a CUDA tensor (None -> NULL).  Refuses CPU tensors: there is no host path."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("manifold_gp_b200: expected a CUDA tensor (no CPU fallback exists)")
    return c_void_p(t.data_ptr())


def stream():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def call(name: str, *args) -> None:
    rc = getattr(_dll, name)(*args)
    if rc != 0:
        raise RuntimeError(f"{name} failed ({rc}): {last_error()}")


MGP_EUNSUPPORTED = -4


def call_rc(name: str, *args) -> int:
    """Like ``call`` but returns the status code instead of raising (for calls with a defined alternative kernel)."""
    return int(getattr(_dll, name)(*args))


def query(name: str, *args) -> int:
    return int(getattr(_dll, name)(*args))


def workspace(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
