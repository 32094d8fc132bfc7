"""Host drivers of the CUDA solvers: batched CG (mBCG), Lanczos, SLQ log-det, and the dispatch rules of
``LinearOperator.solve / inv_quad_logdet / diagonalization``.

These replace the linear_operator loops the reference reaches at utils/train_model.py:55,67-68,
operators/precision_matern_operator.py:53, operators/schur_complement_operator.py:28 and
operators/graph_laplacian_operator.py:132-135 (third party; algorithm restated in oracle/solvers.py).  The
arithmetic of every iteration is in libmgp_b200.so (fused SpMM with dot epilogue + fused vector updates);
the host only enqueues launches and polls one device flag every ``settings.cg_check_interval`` iterations.
"""
from __future__ import annotations

import warnings

import math
import os
import time

import torch

from . import _lib, settings
from ._compat.linear_operator import DenseEigenvectors
from ._lib import c_double, c_float, c_int32, c_int64, ptr, stream

# state layout (must match csrc/cg.cu)
# chunks of CG iterations kept enqueued ahead of the host's convergence check: the GPU rides out host hiccups of up to
# (_FLAG_DEPTH - 1) x check_interval iterations (~12 ms at cfg-C); at most that many iterations run past convergence (the
# vector kernels are no-ops then, the SpMMs are not: ~4 ms per chunk at cfg-C)
_FLAG_DEPTH = 3
S_RHSNORM, S_RZ, S_PAP, S_ALPHA, S_BETA, S_RESID, S_RHSZERO, S_CONV, S_NARR = range(9)
K_MEAN, K_DONE, K_ITER = 0, 1, 2
MAX_CG_COLS = 128


def _pad_ld(c: int, dtype) -> int:
    v = 4 if dtype == torch.float32 else 2
    return (c + v - 1) // v * v


class CGInfo(dict):
    pass


_DIST_BACKEND = None


def set_distributed_backend(backend) -> None:
    """Route the CG solves of covered operators through a multi-GPU backend (``distributed.DistBackend``); ``None`` restores
    the single-GPU drivers.  Must be set identically on every rank of the process group."""
    global _DIST_BACKEND
    _DIST_BACKEND = backend


def _cg_chunk(op, rhs, n_tridiag, tolerance, eps, stop_updating_after, max_iter, n_tridiag_iter):
    """mBCG on <= 128 columns.  Returns (solution [n,C], hist (cpu) or None, info)."""
    if _DIST_BACKEND is not None and _DIST_BACKEND.supports(op, rhs):
        return _DIST_BACKEND.cg_chunk(op, rhs, n_tridiag, tolerance, eps, stop_updating_after, max_iter, n_tridiag_iter)
    n, c = rhs.shape
    dt, dev = rhs.dtype, rhs.device
    sfx = _lib.suffix(dt)
    fused = hasattr(op, "_mgp_matvec") and getattr(op, "_native", lambda: True)()
    ld = _pad_ld(c, dt) if fused else c
    fl = c_float if dt == torch.float32 else c_double

    max_hist = max(1, min(max_iter, n_tridiag_iter)) if n_tridiag else 1
    check = max(1, int(settings.cg_check_interval.value()))
    # Work buffers and the captured graph of `check` iterations are kept on the operator between solves, valid as long as
    # every array the matvec reads is the same memory holding the same version (op._mgp_cache_key): a repeated solve with
    # a new right-hand side (posterior evaluation, bench) replays the graph from the first chunk on.  Capturing costs
    # 5 - 800 ms per solve on B200 boxes (measured), far more than it saves on a 0.6 s solve if it is redone every time.
    ckey = op._mgp_cache_key(dt) if (fused and hasattr(op, "_mgp_cache_key")) else None
    bkey = (n, c, dt, bool(n_tridiag), max_hist, check)
    entry = None
    if ckey is not None:
        cache = getattr(op, "_mgp_cg_cache", None)
        if cache is None or cache["key"] != ckey:
            cache = {"key": ckey, "entries": {}}
            op._mgp_cg_cache = cache
        entry = cache["entries"].get(bkey)
    if entry is not None:
        x, r, p, v, tmp, state, ws, hist = entry["buffers"]
        state.zero_()
        if hist is not None:
            hist.zero_()
    else:
        x = torch.empty((n, ld), dtype=dt, device=dev)
        r = torch.empty((n, ld), dtype=dt, device=dev)
        p = torch.empty((n, ld), dtype=dt, device=dev)
        v = torch.zeros((n, ld), dtype=dt, device=dev)
        tmp = torch.zeros((n, ld), dtype=dt, device=dev) if fused else None
        state = torch.zeros(_lib.query("mgp_cg_state_elems", c_int32(c)), dtype=dt, device=dev)
        ws = torch.zeros(_lib.query("mgp_cg_ws_bytes", c_int64(n), c_int32(c)), dtype=torch.uint8, device=dev)
        hist = torch.zeros((max_hist, 2, c), dtype=dt, device=dev) if n_tridiag else None
        if ckey is not None:
            entry = {"buffers": (x, r, p, v, tmp, state, ws, hist), "graph": None, "launches": 0}
            if len(cache["entries"]) >= 2:          # bound the memory kept alive per operator
                cache["entries"].pop(next(iter(cache["entries"])))
            cache["entries"][bkey] = entry
    st = op._mgp_structure() if fused else None     # fused path runs in the structure's (permuted) row order
    rhs_c = st.to_internal(rhs) if st is not None else rhs
    rhs_c = rhs_c if rhs_c.stride(1) == 1 else rhs_c.contiguous()

    _lib.call("mgp_cg_init_" + sfx, ptr(rhs_c), c_int64(rhs_c.stride(0)), ptr(x), ptr(r), ptr(p), c_int64(ld), c_int64(n),
              c_int32(c), fl(tolerance), fl(eps), fl(stop_updating_after), c_int32(max_iter),
              c_int32(n_tridiag_iter if n_tridiag else 0), ptr(state), ptr(ws), stream())
    pap = state[S_PAP * c:(S_PAP + 1) * c]
    hist_p, hist_n = ptr(hist), c_int32(max_hist if n_tridiag else 0)
    done_flag = state[S_NARR * c + K_DONE:]     # the SpMM launches of chunks replayed past convergence are no-ops too

    def iteration():
        # matvec (p^T A p out of the last SpMM launch) -> r -= alpha v, scalars -> x += alpha p, p = r + beta p
        if fused:
            op._mgp_matvec(p, v, tmp, dot_with=p, dot_out=pap, ncols=c, done_flag=done_flag)
            have_pap = 1
        else:
            with torch.no_grad():
                out = op._matmul(p)
            v.copy_(out)
            have_pap = 0
        _lib.call("mgp_cg_alpha_" + sfx, ptr(p), ptr(v), c_int64(ld), c_int64(n), c_int32(c), c_int32(have_pap),
                  ptr(state), ptr(ws), stream())
        _lib.call("mgp_cg_rupdate_" + sfx, ptr(r), ptr(v), c_int64(ld), c_int64(n), c_int32(c), ptr(state), hist_p, hist_n,
                  None, ptr(ws), stream())
        _lib.call("mgp_cg_pxupdate_" + sfx, ptr(x), ptr(p), ptr(r), c_int64(ld), c_int64(n), c_int32(c), ptr(state), stream())

    # long solves replay whole chunks from a CUDA graph: an iteration is 4 launches of 30-150 us each and the launch gaps
    # + ctypes calls cost ~5 % of it.  Every kernel is a no-op once the done flag is set, so replaying a full chunk past
    # convergence is harmless.  The first chunk runs eagerly (attributes set, value layouts cached, short solves unaffected).
    use_graph = fused and settings.cg_cuda_graph.on() and max_iter >= 4 * check
    cuda_graph, launches_per_replay = None, 0
    if entry is not None and entry["graph"] is not None and use_graph:
        cuda_graph, launches_per_replay = entry["graph"], entry["launches"]
    k = 0
    done = 0.0
    scal = S_NARR * c
    _dbg = os.environ.get("MGP_CG_TIMING") is not None
    _t = {"start": time.perf_counter(), "capture": 0.0, "max_wait": 0.0, "waits": 0} if _dbg else None
    # The done flag of chunk j is read while chunk j + 1 is already enqueued (pinned-memory copy + event): the GPU never
    # waits for the host between chunks.  Running one chunk past convergence is harmless (every kernel is a no-op then).
    flags = torch.zeros(_FLAG_DEPTH, dtype=dt).pin_memory()
    events = [torch.cuda.Event() for _ in range(_FLAG_DEPTH)]
    pending = []          # slots whose flag copy is in flight, oldest first
    slot = 0
    while k < max_iter:
        steps = min(check, max_iter - k)
        if cuda_graph is not None and steps == check:
            cuda_graph.replay()
            _lib._dll.mgp_add_launch_count(launches_per_replay)
        else:
            for _ in range(steps):
                iteration()
        k += steps
        flags[slot:slot + 1].copy_(state[scal + K_DONE:scal + K_DONE + 1], non_blocking=True)
        events[slot].record()
        pending.append(slot)
        slot = (slot + 1) % _FLAG_DEPTH
        if cuda_graph is None:
            # eager chunks (short solves, the first chunk): check at once, then capture the graph for the rest
            s0 = pending.pop(0)
            events[s0].synchronize()
            done = float(flags[s0])
            if done != 0.0:
                break
            if use_graph and max_iter - k >= check:
                _tc = time.perf_counter()
                before = _lib.launch_count()
                # the captured launches bake in raw pointers to the value layouts the structure hands out (st.wi_values /
                # padded_values: a 4-deep LRU owned by the structure).  The cache entry must own them, or a layout evicted by
                # other bandwidths could be freed / recycled under a graph that still replays it.
                if st is not None:
                    st.__dict__["_layout_log"] = []
                try:
                    cuda_graph = _capture(iteration, check)
                finally:
                    keep = st.__dict__.pop("_layout_log", None) if st is not None else None
                launches_per_replay = _lib.launch_count() - before        # kernels of ours inside one replay
                _lib._dll.mgp_add_launch_count(-launches_per_replay)      # the capture pass itself executed nothing
                if entry is not None:
                    entry["graph"], entry["launches"], entry["keepalive"] = cuda_graph, launches_per_replay, keep
                if _dbg:
                    _t["capture"] = time.perf_counter() - _tc
        elif len(pending) == _FLAG_DEPTH:
            s0 = pending.pop(0)
            _tw = time.perf_counter()
            events[s0].synchronize()
            if _dbg:
                _w = time.perf_counter() - _tw
                _t["max_wait"] = max(_t["max_wait"], _w); _t["waits"] += 1
            done = float(flags[s0])
            if done != 0.0:
                break
    if done == 0.0:
        for s0 in pending:
            events[s0].synchronize()
            done = float(flags[s0])
            if done != 0.0:
                break
    if _dbg:
        torch.cuda.synchronize()
        _t["total"] = time.perf_counter() - _t["start"]
        import sys as _sys
        print("[cg timing] " + " ".join(f"{k_}={v_:.4f}" if isinstance(v_, float) else f"{k_}={v_}" for k_, v_ in _t.items() if k_ != "start"),
              file=_sys.stderr)
    out = torch.empty((n, c), dtype=dt, device=dev)
    _lib.call("mgp_cg_finalize_" + sfx, ptr(x), c_int64(ld), ptr(out), c_int64(c), c_int64(n), c_int32(c), ptr(state), stream())
    if st is not None:
        out = st.to_external(out)
    tail = state[scal:scal + 3].tolist()
    info = CGInfo(iterations=int(tail[K_ITER]), mean_residual=float(tail[K_MEAN]), converged=(done == 1.0),
                  residual_norm=state[S_RESID * c:(S_RESID + 1) * c].clone())
    return out, (hist.cpu() if n_tridiag else None), info


_CAPTURE_STREAM = None


def _capture(fn, reps):
    """Capture ``reps`` calls of ``fn`` into a CUDA graph.  Deliberately NOT ``with torch.cuda.graph(...)``: that context
    manager runs ``gc.collect()`` and ``torch.cuda.empty_cache()`` on entry, which was measured to take anywhere from 5 ms
    to 800 ms per capture on B200 boxes (per-solve times of 0.64 s .. 1.45 s for the same 0.63 s of GPU work).  A side
    stream and capture_begin / capture_end only."""
    global _CAPTURE_STREAM
    if _CAPTURE_STREAM is None:
        _CAPTURE_STREAM = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    cur = torch.cuda.current_stream()
    _CAPTURE_STREAM.wait_stream(cur)
    with torch.cuda.stream(_CAPTURE_STREAM):
        g.capture_begin()
        try:
            for _ in range(reps):
                fn()
        finally:
            g.capture_end()
    cur.wait_stream(_CAPTURE_STREAM)
    return g


def _tridiag_from_hist(hist, n_rows, n_tridiag, dtype):
    """Lanczos tridiagonals from the CG coefficients, as linear_cg assembles them (incl. the update_tridiag latch)."""
    t = torch.zeros(n_rows, n_rows, n_tridiag, dtype=hist.dtype)
    update = True
    last = 0
    prev_ar = prev_b = None
    for k in range(n_rows):
        if not update:
            break
        a = hist[k, 0, :n_tridiag].clone()
        b = hist[k, 1, :n_tridiag].clone()
        ar = a.masked_fill(a.eq(0), 1).reciprocal()
        if k == 0:
            t[0, 0] = ar
        else:
            t[k, k] = ar + prev_b * prev_ar
            off = prev_b.sqrt() * prev_ar
            t[k, k - 1] = off
            t[k - 1, k] = off
            if t[k - 1, k].max() < 1e-6:
                update = False
        last = k
        prev_ar, prev_b = ar, b
    t = t[: last + 1, : last + 1]
    return t.permute(2, 0, 1).contiguous().to(dtype)


_in_polish = [False]


def _polish(op, rhs, x, tolerance, max_iter, eps, stop_updating_after):
    """x + d with A d = b - A x (see ``settings.cg_polish``).  The correction solve runs to ``8 * tolerance * |b| / |r|`` per the
    published mean-over-columns rule, i.e. until the correction's recurrence residual is 8 x ``tolerance`` relative to b: the
    fp32 evaluation of b - A x itself has a floor of ~7e-6 at cfg-C (eps_32 * |A| * |x| / |b|), so asking for more buys nothing
    (measured: 132 extra iterations to the full tolerance and the same 7.4e-6 true residual)."""
    with torch.no_grad():
        r = rhs - op._matmul(x)
        bn = rhs.norm(dim=0).clamp_min(torch.finfo(rhs.dtype).tiny)
        rel = r.norm(dim=0) / bn
        before = float(rel.mean())
        out = {"true_residual_before": before, "iterations": 0}
        if not (before > 8.0 * tolerance) or not math.isfinite(before):
            return x, out
        _in_polish[0] = True
        try:
            with warnings.catch_warnings():
                warnings.simplefilter("ignore", RuntimeWarning)
                d, dinfo = linear_cg(op, r, tolerance=min(0.5, 8.0 * tolerance / before), eps=eps, stop_updating_after=stop_updating_after,
                                     max_iter=max(1, min(max_iter, 200)), max_tridiag_iter=0, return_info=True)
        finally:
            _in_polish[0] = False
        out["iterations"] = int(dinfo["iterations"])
        return x + d, out


def linear_cg(op, rhs, n_tridiag=0, tolerance=None, eps=1e-10, stop_updating_after=1e-10, max_iter=None,
              max_tridiag_iter=None, return_info=False):
    """CUDA mBCG.  ``op`` is a LinearOperator (``_matmul``; the fused path is used when it offers ``_mgp_matvec``).
    Same arguments / defaults / stopping rules as ``linear_operator.utils.linear_cg`` (no preconditioner)."""
    if not rhs.is_cuda:
        raise RuntimeError("linear_cg: rhs must be a CUDA tensor (no CPU fallback exists)")
    is_vector = rhs.dim() == 1
    if is_vector:
        rhs = rhs.unsqueeze(-1)
    if max_iter is None:
        max_iter = settings.max_cg_iterations.value()
    if max_tridiag_iter is None:
        max_tridiag_iter = settings.max_lanczos_quadrature_iterations.value()
    if tolerance is None:
        tolerance = settings.eval_cg_tolerance.value() if settings._use_eval_tolerance.on() else settings.cg_tolerance.value()
    if max_tridiag_iter > max_iter:
        raise RuntimeError("Getting a tridiagonalization larger than the number of CG iterations run is not possible!")
    n, c = rhs.shape
    n_iter = min(max_iter, n) if settings.terminate_cg_by_size.on() else max_iter
    n_tridiag_iter = min(max_tridiag_iter, n)
    if n_tridiag > MAX_CG_COLS:
        raise RuntimeError(f"linear_cg: at most {MAX_CG_COLS} tridiagonals per call")
    outs, infos, hist0 = [], [], None
    for c0 in range(0, c, MAX_CG_COLS):
        # NB column chunks converge independently (the published mean-over-columns test is applied per chunk)
        nt = n_tridiag if c0 == 0 else 0
        blk = rhs[:, c0:c0 + MAX_CG_COLS]
        cb = blk.shape[1]
        # Column counts the 128-bit SpMM kernels cannot take (not a multiple of 4 fp32 / 2 fp64 columns: e.g. the 10 probe
        # vectors + 1 right-hand side of the training loss) would fall to the un-pipelined tiled kernel (2x slower per
        # SpMM).  Zero columns are free in a 64-byte-row pass, so the block is padded to the next multiple of 16 (8) columns;
        # a zero column has residual exactly 0, hence comparing the padded mean with tolerance * cb / cpad is the published
        # mean-over-columns stopping rule on the original block.
        vecw, roww = (4, 16) if rhs.dtype == torch.float32 else (2, 8)
        cpad = (cb + roww - 1) // roww * roww
        fused = hasattr(op, "_mgp_matvec") and getattr(op, "_native", lambda: True)()
        if fused and cb >= 5 and cb % vecw != 0 and cpad <= MAX_CG_COLS:
            blk_p = torch.zeros((n, cpad), dtype=rhs.dtype, device=rhs.device)
            blk_p[:, :cb] = blk
            o, h, info = _cg_chunk(op, blk_p, nt, float(tolerance) * cb / cpad, float(eps), float(stop_updating_after),
                                   int(n_iter), int(n_tridiag_iter))
            o = o[:, :cb]
            info["mean_residual"] = info["mean_residual"] * cpad / cb
            info["residual_norm"] = info["residual_norm"][:cb]
        else:
            o, h, info = _cg_chunk(op, blk, nt, float(tolerance), float(eps), float(stop_updating_after),
                                   int(n_iter), int(n_tridiag_iter))
        outs.append(o)
        infos.append(info)
        if c0 == 0:
            hist0 = h
    result = outs[0] if len(outs) == 1 else torch.cat(outs, dim=1)
    polish = settings.cg_polish.value()
    want_polish = (polish is True) or (polish == "auto" and rhs.dtype == torch.float32 and float(tolerance) <= 1e-4)
    polish_info = None
    if want_polish and not _in_polish[0] and all(i["converged"] for i in infos) and all(i["iterations"] > 0 for i in infos):
        result, polish_info = _polish(op, rhs, result, float(tolerance), int(n_iter), float(eps), float(stop_updating_after))
    info = infos[0] if len(infos) == 1 else CGInfo(
        iterations=max(i["iterations"] for i in infos), converged=all(i["converged"] for i in infos),
        mean_residual=sum(i["mean_residual"] for i in infos) / len(infos),
        residual_norm=torch.cat([i["residual_norm"] for i in infos]))
    if polish_info is not None:
        info["polish"] = polish_info
    if not info["converged"] and n_iter > 0:
        warnings.warn("CG terminated in {} iterations with average residual norm {} which is larger than the tolerance "
                      "of {} specified by settings.cg_tolerance.".format(info["iterations"], info["mean_residual"], tolerance),
                      RuntimeWarning)
    if is_vector:
        result = result.squeeze(-1)
    if n_tridiag:
        k_done = infos[0]["iterations"]
        rows = k_done - 1 if infos[0]["converged"] else k_done
        rows = max(1, min(rows, n_tridiag_iter))
        t_mat = _tridiag_from_hist(hist0, rows, n_tridiag, rhs.dtype).to(rhs.device)
        return (result, t_mat, info) if return_info else (result, t_mat)
    return (result, info) if return_info else result


# ---- Lanczos ----------------------------------------------------------------------------------------------------------
def lanczos_tridiag(op, max_iter, init_vec=None, tol=1e-5, generator=None):
    """Lanczos with full re-orthogonalisation.  Returns ``q_mat[j, n]`` (vector-major!) and ``t_mat[j, j]``.

    Each step: r = A q_k (fused SpMM), then classical Gram-Schmidt against ALL previous vectors in one fused pass
    (``mgp_lanczos_reorth``: removes alpha_k q_k, beta_{k-1} q_{k-1} and every other component; alpha_k = c[k]),
    then check passes (``mgp_lanczos_dots`` / ``mgp_lanczos_axpy``) while any |q_j . r| / |r| > tol (<= 10), as
    linear_operator.utils.lanczos.lanczos_tridiag does; early exit when beta < 1e-6.
    """
    n = op.shape[0]
    dt, dev = op.dtype, op.device
    if dev.type != "cuda":
        raise RuntimeError("lanczos_tridiag: operator must live on a CUDA device (no CPU fallback exists)")
    sfx = _lib.suffix(dt)
    num_iter = min(int(max_iter), n)
    if init_vec is None:
        init_vec = torch.randn(n, dtype=dt, device=dev, generator=generator)
    q = torch.empty((num_iter + 1, n), dtype=dt, device=dev)
    r = init_vec.reshape(n).to(dt).clone()
    c = torch.zeros(num_iter + 1, dtype=dt, device=dev)
    nrm2 = torch.zeros(1, dtype=dt, device=dev)
    beta = torch.zeros(1, dtype=dt, device=dev)
    ws = torch.zeros(_lib.query("mgp_lanczos_ws_bytes", c_int64(n), c_int32(num_iter + 1)), dtype=torch.uint8, device=dev)
    alphas = torch.zeros(num_iter, dtype=dt, device=dev)
    betas = torch.zeros(num_iter, dtype=dt, device=dev)
    fused = hasattr(op, "_mgp_matvec") and getattr(op, "_native", lambda: True)()
    tmp = torch.empty((n, 1), dtype=dt, device=dev) if fused else None
    st = op._mgp_structure() if fused else None     # fused path: Lanczos vectors live in the structure's row order
    if st is not None:
        r = st.to_internal(r).contiguous()

    _lib.call("mgp_lanczos_reorth_" + sfx, None, c_int64(n), c_int32(0), ptr(r), c_int64(n), ptr(c), ptr(nrm2), ptr(ws), stream())
    _lib.call("mgp_lanczos_normalize_" + sfx, ptr(r), c_int64(n), ptr(nrm2), ptr(q[0]), None, stream())
    steps = 0
    for k in range(num_iter):
        qk = q[k].unsqueeze(-1)
        if fused:
            op._mgp_matvec(qk, r.unsqueeze(-1), tmp, ncols=1)
        else:
            with torch.no_grad():
                r.copy_(op._matmul(qk).reshape(n))
        j = k + 1
        _lib.call("mgp_lanczos_reorth_" + sfx, ptr(q), c_int64(n), c_int32(j), ptr(r), c_int64(n), ptr(c), ptr(nrm2), ptr(ws), stream())
        alphas[k:k + 1].copy_(c[k:k + 1])
        steps = k + 1
        if k + 1 >= num_iter:
            break
        ok = False
        for _ in range(10):
            _lib.call("mgp_lanczos_dots_" + sfx, ptr(q), c_int64(n), c_int32(j), ptr(r), c_int64(n), ptr(c), ptr(ws), stream())
            worst, nr2 = torch.stack((c[:j].abs().max(), nrm2[0])).tolist()   # one device->host read per pass
            if nr2 <= 0.0 or worst / (nr2 ** 0.5) <= tol:
                ok = True
                break
            _lib.call("mgp_lanczos_axpy_" + sfx, ptr(q), c_int64(n), c_int32(j), ptr(r), c_int64(n), ptr(c), ptr(nrm2), ptr(ws), stream())
        _lib.call("mgp_lanczos_normalize_" + sfx, ptr(r), c_int64(n), ptr(nrm2), ptr(q[k + 1]), ptr(beta), stream())
        betas[k:k + 1].copy_(beta)
        if nr2 ** 0.5 <= 1e-6 or not ok:
            break
    t = torch.diag(alphas[:steps])
    if steps > 1:
        off = betas[:steps - 1]
        t = t + torch.diag(off, 1) + torch.diag(off, -1)
    qm = q[:steps]
    if st is not None and st.inv is not None:
        qm = qm.index_select(1, st.inv)
    return qm, t


def lanczos_tridiag_to_diag(t_mat):
    """eigh of the tridiagonal; negative Ritz values -> 1 with zeroed vectors (published behaviour)."""
    evals, evecs = torch.linalg.eigh(t_mat)
    mask = evals.ge(0)
    evecs = evecs * mask.to(evecs.dtype).unsqueeze(-2)
    evals = evals.masked_fill(~mask, 1)
    return evals, evecs


def smallest_eigenpairs(op, num_modes, rtol=1e-6, degree=24, max_outer=60, extra=None, generator=None, return_info=False):
    """The ``num_modes`` smallest eigenpairs of a symmetric positive semi-definite operator by Chebyshev-filtered subspace
    iteration (Zhou & Saad): a block of ``num_modes + extra`` vectors is repeatedly pushed through a degree-``degree``
    Chebyshev polynomial of the operator that damps [largest Ritz value of the block, lambda_max] (``degree`` batched SpMMs,
    every 16 columns one pass of the fused SpMM kernel), re-orthonormalised and Rayleigh-Ritz rotated, until
    ``|A x - theta x| <= rtol * lambda_max`` for the wanted pairs.

    Why it exists: ``RiemannKernel.eval`` needs the SMALLEST eigenpairs (riemann_kernel.py:117-130 gets them from a dense
    ``eigh``, O(N^3)).  ``diagonalization(method="lanczos")`` reproduces linear_operator's 3 * num_modes-step Lanczos, whose
    Ritz values converge at the TOP of the spectrum first: at N >> 3 * num_modes (BASELINE cfg-B: N = 70k, 500 modes) its
    lowest ``num_modes`` Ritz pairs are not eigenpairs (measured residual |L phi - lambda phi| ~ 1).  The filter turns the
    same matvec kernel into a solver that converges at the bottom.  Dense pieces (QR of [N, b], b x b ``eigh``) are plain
    library calls on small matrices."""
    n = op.shape[0]
    dt, dev = op.dtype, op.device
    if dev.type != "cuda":
        raise RuntimeError("smallest_eigenpairs: operator must live on a CUDA device (no CPU fallback exists)")
    m = int(min(num_modes, n))
    if extra is None:
        extra = max(32, m // 4)
    b = min(n, ((m + extra + 15) // 16) * 16)
    with torch.no_grad():
        # upper bound of the spectrum: largest Ritz value of a short Lanczos run + its residual norm
        _, t = lanczos_tridiag(op, min(40, n), generator=generator)
        th = torch.linalg.eigvalsh(t.double())
        lam_max = float(th[-1]) + (float(t[-1, -2].abs()) if t.shape[0] > 1 else 0.0)
        lam_max *= 1.01
        x = torch.randn(n, b, dtype=dt, device=dev, generator=generator)
        x, _ = torch.linalg.qr(x)
        info = {"outer": 0, "matvec_columns": 0, "lam_max": lam_max, "block": b}
        theta = None
        for outer in range(max_outer + 1):
            ax = op._matmul(x)
            info["matvec_columns"] += b
            h = (x.transpose(0, 1) @ ax).double()
            theta, v = torch.linalg.eigh(0.5 * (h + h.transpose(0, 1)))
            v = v.to(dt)
            x, ax = x @ v, ax @ v
            res = (ax - x * theta.to(dt)).norm(dim=0)
            worst = float(res[:m].max())
            info.update(outer=outer, residual=worst)
            if worst <= rtol * lam_max or outer == max_outer:
                break
            # damp [a, lam_max]; a = largest Ritz value of the block, a0 = smallest (the scaling point)
            a, a0 = float(theta[-1]), float(theta[0])
            a = max(a, a0 + 1e-6 * lam_max)
            e, c = 0.5 * (lam_max - a), 0.5 * (lam_max + a)
            sigma = e / (a0 - c)
            sigma1 = sigma
            y = (ax - c * x) * (sigma1 / e)
            for _ in range(2, degree + 1):
                sigma2 = 1.0 / (2.0 / sigma1 - sigma)
                ay = op._matmul(y)
                info["matvec_columns"] += b
                ynew = (ay - c * y) * (2.0 * sigma2 / e) - (sigma * sigma2) * x
                x, y, sigma = y, ynew, sigma2
            x, _ = torch.linalg.qr(y)
        evals = theta[:m].to(dt)
        evecs = x[:, :m].contiguous()
    if return_info:
        return evals, evecs, info
    return evals, evecs


def diagonalization(op, method=None):
    """``LinearOperator.diagonalization``: 'symeig' (dense) when size <= max_cholesky_size, else 'lanczos' with
    ``settings.max_root_decomposition_size`` steps.  Returns (evals ascending, DenseEigenvectors)."""
    n = op.shape[0]
    if method is None:
        method = "symeig" if n <= settings.max_cholesky_size.value() else "lanczos"
    if method == "symeig":
        with torch.no_grad():
            evals, evecs = torch.linalg.eigh(op.to_dense())
        return evals, DenseEigenvectors(evecs)
    if method == "lanczos":
        be = _DIST_BACKEND
        if be is not None and getattr(be, "world", 1) > 1 and n >= be.min_rows and hasattr(op, "structure") and \
                getattr(op, "normalization", None) == "symmetric" and not getattr(op, "transposed", False):
            with torch.no_grad():          # row-partitioned Lanczos over the process group (distributed.dist_lanczos_tridiag)
                evals, evecs, _ = be.lanczos_eigenpairs(op, settings.max_root_decomposition_size.value())
            return evals, DenseEigenvectors(evecs)
        with torch.no_grad():
            q, t = lanczos_tridiag(op, settings.max_root_decomposition_size.value())
            evals, v = lanczos_tridiag_to_diag(t)
            evecs = q.T @ v       # plain library GEMM [n, j] x [j, j]
        return evals, DenseEigenvectors(evecs)
    if method == "chebyshev":
        evals, evecs = smallest_eigenpairs(op, min(n, max(1, settings.max_root_decomposition_size.value() // 3)))
        return evals, DenseEigenvectors(evecs)
    raise RuntimeError(f"Unknown diagonalization method '{method}'")


# ---- solve / inv_quad_logdet ---------------------------------------------------------------------------------------------
def _params_requiring_grad(op):
    return [t for t in op.representation() if t.is_floating_point() and t.requires_grad]


def _cg_tol():
    return settings.eval_cg_tolerance.value() if settings._use_eval_tolerance.on() else settings.cg_tolerance.value()


class _CGSolveFn(torch.autograd.Function):
    """x = A^-1 b by CUDA CG; backward: g_b = A^-1 g, g_theta = -(A^-1 g)^T (dA/dtheta) x  (A symmetric)."""

    @staticmethod
    def forward(ctx, op, rhs, *params):
        ctx.op = op
        with torch.no_grad():
            with warnings.catch_warnings():
                warnings.simplefilter("ignore", RuntimeWarning)
                x = linear_cg(op, rhs, tolerance=_cg_tol())
        ctx.save_for_backward(x)
        ctx.n_params = len(params)
        return x

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        op = ctx.op
        with torch.no_grad():
            with warnings.catch_warnings():
                warnings.simplefilter("ignore", RuntimeWarning)
                u = linear_cg(op, g.contiguous(), tolerance=_cg_tol())
        grads = [None] * ctx.n_params
        params = _params_requiring_grad(op)
        if params:
            with torch.enable_grad():
                val = -(u * op._matmul(x)).sum()
                pg = torch.autograd.grad(val, params, allow_unused=True)
            # map back onto the positional params given to forward (same order as _params_requiring_grad)
            grads = list(pg)
        return (None, u if ctx.needs_input_grad[1] else None, *grads)


def solve(op, rhs):
    """``LinearOperator.solve``: dense Cholesky when size <= max_cholesky_size, CUDA CG otherwise."""
    n = op.shape[0]
    squeeze = rhs.dim() == 1
    r = rhs.unsqueeze(-1) if squeeze else rhs
    if n <= settings.max_cholesky_size.value():
        chol = torch.linalg.cholesky(op.to_dense())
        out = torch.cholesky_solve(r, chol)
    else:
        params = _params_requiring_grad(op) if torch.is_grad_enabled() else []
        if params or (torch.is_grad_enabled() and r.requires_grad):
            out = _CGSolveFn.apply(op, r, *params)
        else:
            with warnings.catch_warnings():
                warnings.simplefilter("ignore", RuntimeWarning)
                out = linear_cg(op, r, tolerance=_cg_tol())
    return out.squeeze(-1) if squeeze else out


def inv_quad_logdet(op, inv_quad_rhs=None, logdet=False, reduce_inv_quad=True, probes=None, return_info=False):
    """``LinearOperator.inv_quad_logdet``.

    size <= max_cholesky_size: dense Cholesky (differentiable through the dense matmul).
    otherwise: one mBCG run on [probes | rhs]; log|A| by stochastic Lanczos quadrature from the CG tridiagonals;
    gradients via the same identities linear_operator's ``InvQuadLogdet.backward`` uses, expressed as surrogates
    through one differentiable operator matvec:  d log|A| = E_z[(A^-1 z)^T dA z],  d (r^T A^-1 r) = -s^T dA s + 2 s^T dr.
    """
    n = op.shape[0]
    dt, dev = op.dtype, op.device
    zero = torch.zeros((), dtype=dt, device=dev)
    rhs = None
    if inv_quad_rhs is not None:
        rhs = inv_quad_rhs if inv_quad_rhs.dim() == 2 else inv_quad_rhs.unsqueeze(-1)
    if n <= settings.max_cholesky_size.value():
        chol = torch.linalg.cholesky(op.to_dense())
        iq = zero
        if rhs is not None:
            sol = torch.linalg.solve_triangular(chol, rhs.to(dt), upper=False)
            iq = sol.pow(2).sum(-2)
            if reduce_inv_quad:
                iq = iq.sum(-1)
        ld = chol.diagonal().log().sum() * 2 if logdet else zero
        return (iq, ld, None) if return_info else (iq, ld)

    cols = []
    n_probe = 0
    if logdet:
        if probes is None:
            probes = torch.randn(n, settings.num_trace_samples.value(), dtype=dt, device=dev)
        pn = probes.norm(2, dim=-2, keepdim=True)
        probes_unit = probes / pn
        cols.append(probes_unit)
        n_probe = probes.shape[1]
    if rhs is not None:
        cols.append(rhs.detach().to(dt))
    allrhs = torch.cat(cols, dim=1).contiguous()
    with torch.no_grad():
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", RuntimeWarning)
            if n_probe:
                solves, t_mat, info = linear_cg(op, allrhs, n_tridiag=n_probe, tolerance=_cg_tol(), return_info=True)
            else:
                solves, info = linear_cg(op, allrhs, tolerance=_cg_tol(), return_info=True)
    iq, ld = zero, zero
    need_grad = torch.is_grad_enabled() and (bool(_params_requiring_grad(op)) or (rhs is not None and inv_quad_rhs.requires_grad))
    if logdet:
        evals, evecs = lanczos_tridiag_to_diag(t_mat)
        ld = (n / float(n_probe)) * (evecs[..., 0, :].pow(2) * evals.log()).sum(-1).sum(0)
    if rhs is not None:
        s = solves[:, n_probe:]
        iq = (s * rhs.detach()).sum(-2)
    if need_grad:
        # one differentiable matvec on [z | s]
        zs = []
        if logdet:
            zs.append(probes_unit * pn)           # un-normalised probes z
        if rhs is not None:
            zs.append(solves[:, n_probe:])
        az = op._matmul(torch.cat(zs, dim=1).contiguous())
        if logdet:
            u = solves[:, :n_probe] * pn          # A^-1 z
            sur = (u * az[:, :n_probe]).sum() / n_probe
            ld = ld + (sur - sur.detach())
        if rhs is not None:
            s = solves[:, n_probe:]
            a_s = az[:, n_probe if logdet else 0:]
            sur = 2.0 * (s * (inv_quad_rhs if inv_quad_rhs.dim() == 2 else inv_quad_rhs.unsqueeze(-1))).sum(-2) - (s * a_s).sum(-2)
            iq = iq + (sur - sur.detach())
    if rhs is not None and reduce_inv_quad:
        iq = iq.sum(-1)
    return (iq, ld, info) if return_info else (iq, ld)
