"""Row-partitioned multi-GPU execution of the Matern-precision CG solve.

The reference has no distributed code; this is new design.  One process per GPU (torchrun), ``torch.distributed`` (NCCL
over NVLink/NVSwitch; gloo in the CPU tests) for the plumbing:

* rows are partitioned into contiguous, 128-aligned blocks of the space-filling-curve order, so a rank's neighbours are
  almost all its own rows and the halo is the thin boundary of a 2-D patch;
* every matvec exchanges ONLY halo rows of the vector (``all_to_all_single`` with precomputed send lists; the receive
  buffer is the halo region of the local vector, ordered by owner so each peer's rows land contiguously);
* every CG inner product is an all-reduce of ``C`` floats (p^T A p comes out of the SpMM epilogue, r^T r out of the fused
  update kernel; ``mgp_cg_dist_scalars`` finishes alpha/beta/flags on every rank identically);
* graph construction: ``PartitionedGraph`` / ``PartitionedPrecision`` (end of this file) search, symmetrise and build structure
  and values per rank -- no rank holds the whole edge list (symmetrise = one all_to_all_v of triples, value build = three local
  passes with two halo gathers); ``bench.py --gpus N`` runs on them.  The round-1 form is kept for ``DistBackend`` (the
  reference's UNCHANGED training loop builds its operators on every rank by construction): kNN queries sharded and
  all-gathered, the O(nnz) structure / value builds replicated, the local operator a *view* of the global tile-compacted
  structure (``LocalStructure``: no index rebasing, row pointers keep global positions so the TMA copies stay 16-byte aligned)
  with the halo column ids translated to local vector rows.
"""
from __future__ import annotations

import json
import os
import time
import warnings

import torch
import torch.distributed as dist


class RowPartition:
    """Contiguous row blocks, aligned to ``align`` rows (the SpMM tile height) except for the last block."""

    def __init__(self, n: int, world: int, align: int = 128):
        self.n, self.world = int(n), int(world)
        per = (n + world - 1) // world
        per = (per + align - 1) // align * align
        self.bounds = [min(r * per, n) for r in range(world + 1)]
        self.bounds[-1] = n

    def range(self, rank: int):
        return self.bounds[rank], self.bounds[rank + 1]

    def owner(self, ids: torch.Tensor) -> torch.Tensor:
        b = torch.tensor(self.bounds[1:], device=ids.device, dtype=ids.dtype)
        return torch.searchsorted(b, ids, right=True)


class HaloPlan:
    """Which rows each rank sends to / receives from every peer for one matvec.

    ``halo_ids``  -- sorted global ids of the out-of-range columns my rows reference (grouped by owner because owners are
                     contiguous ranges); vector row ``n_loc + i`` holds the value of global row ``halo_ids[i]``.
    ``send_idx``  -- local row indices to pack for the peers, concatenated in peer order.
    """

    def __init__(self, part: RowPartition, rank: int, cols_of_my_rows: torch.Tensor, group=None):
        lo, hi = part.range(rank)
        dev = cols_of_my_rows.device
        outside = (cols_of_my_rows < lo) | (cols_of_my_rows >= hi)
        self.halo_ids = torch.unique(cols_of_my_rows[outside].to(torch.int64), sorted=True)
        self.n_loc = hi - lo
        self.lo, self.hi = lo, hi
        owners = part.owner(self.halo_ids)
        self.recv_counts = torch.bincount(owners, minlength=part.world).tolist()
        # tell every peer which of its rows I need
        send_counts_t = torch.zeros(part.world, dtype=torch.int64, device=dev)
        dist.all_to_all_single(send_counts_t, torch.tensor(self.recv_counts, dtype=torch.int64, device=dev), group=group)
        self.send_counts = send_counts_t.tolist()
        req = torch.empty(int(sum(self.send_counts)), dtype=torch.int64, device=dev)
        dist.all_to_all_single(req, self.halo_ids, self.send_counts, self.recv_counts, group=group)
        self.part = part
        self.send_idx = (req - lo).contiguous()          # local row index of every row I must send
        assert self.send_idx.numel() == 0 or (int(self.send_idx.min()) >= 0 and int(self.send_idx.max()) < self.n_loc)
        self.group = group

    def to_local(self, ids: torch.Tensor) -> torch.Tensor:
        """global (permuted) row id -> row of the local vector [own rows | halo rows]."""
        ids = ids.to(torch.int64)
        inside = (ids >= self.lo) & (ids < self.hi)
        pos = torch.searchsorted(self.halo_ids, ids.clamp(min=0))
        pos = pos.clamp(max=max(self.halo_ids.numel() - 1, 0))
        return torch.where(inside, ids - self.lo, self.n_loc + pos)

    def exchange(self, x: torch.Tensor, ncols: int = None) -> None:
        """Fill rows [n_loc, n_loc + H) of ``x`` ([n_loc + H, ld]) with the peers' current values."""
        if self.halo_ids.numel() == 0 and self.send_idx.numel() == 0:
            return
        send = x.index_select(0, self.send_idx)
        dist.all_to_all_single(x[self.n_loc:self.n_loc + self.halo_ids.numel()], send, self.recv_counts, self.send_counts,
                               group=self.group)


class LocalStructure:
    """Duck-types ``graph.GraphStructure`` for the rows of one rank (views of the global arrays + remapped column ids)."""

    def __init__(self, gst, plan: HaloPlan):
        from . import graph
        self.n = plan.n_loc
        self.nnz = gst.nnz
        self.m = gst.m
        self.device = gst.device
        self.perm = self.inv = self.perm32 = None
        lo, hi = plan.lo, plan.hi
        self.rowptr = gst.rowptr[lo:hi + 1]                      # global positions: a / lcol are passed unsliced
        self.col = plan.to_local(gst.col).to(torch.int32).contiguous()
        self._dot_ws = None
        self.tiles = None
        t = gst.tiles
        if t is not None and lo % t["rows"] == 0:
            t0, t1 = lo // t["rows"], (hi + t["rows"] - 1) // t["rows"]
            self.tiles = dict(lcol=t["lcol"], halo_ptr=t["halo_ptr"][t0:t1 + 1],
                              halo_col=plan.to_local(t["halo_col"]).to(torch.int32).contiguous(),
                              lmax=t["lmax"], nzmax=t["nzmax"], rows=t["rows"], halo_total=t["halo_total"])
            if "prowptr" in t:   # padded streams of the pipelined kernel: global positions again, arrays unsliced
                self.tiles.update(prowptr=t["prowptr"][lo:hi + 1], plcol=t["plcol"], nnzp=t["nnzp"], pnzmax=t["pnzmax"])
            if "wptr" in t:      # warp-interleaved streams (v5): 16 blocks per tile, global stream positions, streams unsliced;
                # the metadata arrays are copied (the kernel bulk-copies them in 32-tile chunks: alignment + padding)
                nt = t1 - t0
                wloc = torch.full((512 * ((nt + 31) // 32) + 4,), int(t["wptr"][16 * t1]), dtype=torch.int32, device=self.device)
                wloc[:16 * nt + 1] = t["wptr"][16 * t0:16 * t1 + 1]
                hloc = torch.full((32 * ((nt + 31) // 32) + 4,), int(t["hptr"][t1]), dtype=torch.int32, device=self.device)
                hloc[:nt + 1] = t["hptr"][t0:t1 + 1]
                self.tiles.update(wptr=wloc, wcol=t["wcol"], nnzw=t["nnzw"], wnzmax=t["wnzmax"], hptr=hloc,
                                  hcol=plan.to_local(t["hcol"]).to(torch.int32).contiguous(), hmax=t["hmax"])
                # peer-memory variant: (owner rank << 26) | row inside the owner's block -- the kernel reads the owner's
                # vector directly over NVLink (ids of rows this rank owns are never used by its own tiles' halo lists)
                gid = t["hcol"].to(torch.int64)
                own = plan.part.owner(gid).clamp_(max=plan.part.world - 1)
                starts = torch.tensor(plan.part.bounds[:-1], dtype=torch.int64, device=self.device)
                assert max(b1 - b0 for b0, b1 in zip(plan.part.bounds[:-1], plan.part.bounds[1:])) < (1 << 26)
                self.tiles["hcol_peer"] = ((own << 26) | (gid - starts[own])).to(torch.int32).contiguous()
        self._gst = gst
        self.TILED_SMEM_LIMIT = gst.TILED_SMEM_LIMIT

    def build_tiles(self):
        return self.tiles

    def tiled_ok(self, dtype, cw):
        return self.tiles is not None and self._gst.tiled_ok(dtype, cw)

    def dot_ws(self):
        if self._dot_ws is None:
            self._dot_ws = torch.zeros_like(self._gst.dot_ws())
        return self._dot_ws

    def padded_values(self, a):
        return self._gst.padded_values(a)

    def wi_values(self, a):
        if getattr(self, "_aw_persistent", None) is not None:
            return self._aw_persistent     # DistPrecision.update_values keeps one buffer alive across bandwidths (CUDA graphs)
        return self._gst.wi_values(a)    # built from the GLOBAL rowptr / wptr: stream positions are global


class DistPrecision:
    """(2nu/kappa^2 + L_sym)^nu on the rows of this rank, optionally inside the reference's Scale / Noise wrappers
    (riemann_gp.py:32-39: Noise(Scale(P)) is the operator of the training loss).  ``values`` = (diag[n], a[nnz]) of the GLOBAL
    structure.  With ``coef`` (device scalar c = outputscale or its reciprocal) the operator is Q = c P; with ``noise``
    (device scalar s) it is Q - s Q^2 + s^2 Q^3 = Q (x - s Q (x - s Q x)), i.e. 3 nu chained SpMM launches whose ``x - s Q(.)``
    combinations ride in the launches' epilogues (mgp_wi_ext.ep_coef / ep_add) -- no elementwise kernel between them.

    The value arrays live in buffers owned by this object (``update_values`` copies a new bandwidth's values into them), so the
    solvers' captured CUDA graphs and peer-memory registrations survive a training step."""

    def __init__(self, gst, diag, a, shift, nu: int, part: RowPartition, rank: int, group=None, coef=None, noise=None):
        lo, hi = part.range(rank)
        p0, p1 = int(gst.rowptr[lo]), int(gst.rowptr[hi])
        self.plan = HaloPlan(part, rank, gst.col[p0:p1], group=group)
        self.st = LocalStructure(gst, self.plan)
        self.gst, self.lo, self.hi = gst, lo, hi
        self.nu = nu
        self.n_loc = hi - lo
        self.n_ext = self.n_loc + int(self.plan.halo_ids.numel())
        dt, dev = a.dtype, a.device
        t = gst.tiles
        self.diag = torch.empty(self.n_loc, dtype=dt, device=dev)
        self.shift = torch.empty(1, dtype=dt, device=dev)
        self.coef = torch.ones(1, dtype=dt, device=dev) if (coef is not None or noise is not None) else None
        self.ncoef = torch.zeros(1, dtype=dt, device=dev) if noise is not None else None      # -noise * coef
        self.has_noise = noise is not None
        if t is not None and "wptr" in t:
            self.st._aw_persistent = torch.zeros(t["nnzw"] + 64, dtype=dt, device=dev)
        self.update_values(diag, a, shift, coef, noise)

    @property
    def nstage(self):
        return self.nu * (3 if self.has_noise else 1)

    def update_values(self, diag, a, shift, coef=None, noise=None):
        """New bandwidth / lengthscale / scales: same memory, new contents (nothing the kernels point at moves)."""
        with torch.no_grad():
            self.a = a
            self.diag.copy_(diag[self.lo:self.hi])
            self.shift.copy_(shift.detach().reshape(-1)[:1].to(self.shift.dtype))
            if getattr(self.st, "_aw_persistent", None) is not None:
                self.st._aw_persistent.copy_(self.gst.wi_values(a))
            if self.coef is not None:
                c = torch.ones(1, dtype=self.coef.dtype, device=self.coef.device) if coef is None else \
                    coef.detach().reshape(-1)[:1].to(self.coef.dtype)
                self.coef.copy_(c)
                if self.ncoef is not None:
                    self.ncoef.copy_(-noise.detach().reshape(-1)[:1].to(self.coef.dtype) * c)

    def stage_epilogues(self):
        """Per SpMM stage: (ep_coef tensor or None, add_source: bool).  Stage i is the (i % nu)-th Laplacian step of the
        (i // nu)-th application of Q; the last step of an application carries the wrapper algebra."""
        out = []
        for i in range(self.nstage):
            end_of_q = (i % self.nu) == self.nu - 1
            q = i // self.nu
            if not end_of_q or self.coef is None:
                out.append((None, False))
            elif self.has_noise and q < 2:
                out.append((self.ncoef, True))          # x - s c P(.)
            else:
                out.append((self.coef, False))          # c P(.)
        return out

    def matvec(self, p, out, tmp, ncols, dot_out=None):
        """out[:n_loc, :ncols] <- P p  (p, tmp: [n_ext, ld] with halo rows; out: [>= n_loc, ld]).  NCCL transport: bare
        precision operator only."""
        from . import graph
        if self.coef is not None:
            raise RuntimeError("DistPrecision.matvec: the NCCL transport covers the bare precision operator; wrappers need PeerCG(mode='cg1')")
        src = p
        for s in range(self.nu):
            last = s == self.nu - 1
            dst = out if last else tmp
            self.plan.exchange(src)
            graph.lap_spmm(self.st, self.a, self.diag, src[:, :ncols], shift=self.shift, out=dst[:self.n_loc, :ncols],
                           dot_with=p if last else None, dot_out=dot_out if last else None)
            src = dst
        return out


class DistCG:
    """mBCG on row-partitioned vectors with persistent buffers and a cached CUDA graph of ``check_interval`` iterations.

    Same stopping rules as ``solvers.linear_cg``.  All kernels are no-ops once the done flag is set, so whole chunks can
    be replayed without overshooting; the host reads one flag per chunk.  At 8 GPUs an iteration is ~15 launches of a
    few microseconds each -- without the graph the host, not the GPU, would be the bottleneck."""

    def __init__(self, op: DistPrecision, ncols: int, dtype=torch.float32, tolerance=1e-6, eps=1e-10,
                 stop_updating_after=1e-10, max_iter=1000, check_interval=16, group=None, use_cuda_graph=True):
        from . import _lib, solvers
        from ._lib import c_int32, c_int64
        self.op, self.c, self.dt, self.group = op, ncols, dtype, group
        self.tol, self.eps, self.stop, self.max_iter, self.check = tolerance, eps, stop_updating_after, max_iter, check_interval
        dev = op.a.device
        self.dev = dev
        n_loc = op.n_loc
        ld = solvers._pad_ld(ncols, dtype)
        self.ld = ld
        z = lambda rows: torch.zeros((rows, ld), dtype=dtype, device=dev)
        self.b, self.x, self.r, self.v = z(n_loc), z(n_loc), z(n_loc), z(n_loc)
        self.p, self.tmp = z(op.n_ext), z(op.n_ext)
        self.state = torch.zeros(_lib.query("mgp_cg_state_elems", c_int32(ncols)), dtype=dtype, device=dev)
        self.ws = torch.zeros(_lib.query("mgp_cg_ws_bytes", c_int64(n_loc), c_int32(ncols)), dtype=torch.uint8, device=dev)
        self.rbuf = torch.zeros(ncols, dtype=dtype, device=dev)
        self.pap = self.state[solvers.S_PAP * ncols:(solvers.S_PAP + 1) * ncols]
        self.graph = None
        self.use_graph = use_cuda_graph and dev.type == "cuda"

    def _allreduce_rbuf(self):
        dist.all_reduce(self.rbuf, group=self.group)

    def _after_init(self):
        pass

    def _scalars(self, what):
        from . import _lib
        from ._lib import c_double, c_float, c_int32, ptr, stream
        fl = c_float if self.dt == torch.float32 else c_double
        _lib.call("mgp_cg_dist_scalars_" + _lib.suffix(self.dt), ptr(self.state), ptr(self.rbuf), c_int32(self.c), c_int32(what),
                  fl(self.tol), fl(self.eps), fl(self.stop), c_int32(self.max_iter), c_int32(0), None, c_int32(0), stream())

    def _iteration(self):
        from . import _lib
        from ._lib import c_int32, c_int64, ptr, stream
        sfx = _lib.suffix(self.dt)
        n_loc, c, ld = self.op.n_loc, self.c, self.ld
        self.op.matvec(self.p, self.v, self.tmp, c, dot_out=self.pap)
        dist.all_reduce(self.pap, group=self.group)
        _lib.call("mgp_cg_dist_update_" + sfx, ptr(self.x), ptr(self.r), ptr(self.p), ptr(self.v), c_int64(ld), c_int64(n_loc),
                  c_int32(c), ptr(self.state), ptr(self.rbuf), ptr(self.ws), stream())
        dist.all_reduce(self.rbuf, group=self.group)
        self._scalars(2)
        _lib.call("mgp_cg_pupdate_" + sfx, ptr(self.p), ptr(self.r), c_int64(ld), c_int64(n_loc), c_int32(c), ptr(self.state), stream())

    def solve(self, b_loc: torch.Tensor):
        """``b_loc`` [n_loc, C]: this rank's block of the right-hand sides (structure order).  Returns (x_loc, info)."""
        from . import _lib, solvers
        from ._lib import c_int32, c_int64, ptr, stream
        sfx = _lib.suffix(self.dt)
        n_loc, c, ld = self.op.n_loc, self.c, self.ld
        self.b[:, :c].copy_(b_loc, non_blocking=True)
        self.state.zero_()
        _lib.call("mgp_cg_dist_norm2_" + sfx, ptr(self.b), c_int64(ld), c_int64(n_loc), c_int32(c), ptr(self.state), ptr(self.rbuf),
                  ptr(self.ws), stream())
        self._allreduce_rbuf()
        self._scalars(0)
        _lib.call("mgp_cg_dist_init_" + sfx, ptr(self.b), c_int64(ld), ptr(self.x), ptr(self.r), ptr(self.p), c_int64(ld),
                  c_int64(n_loc), c_int32(c), ptr(self.state), ptr(self.rbuf), ptr(self.ws), stream())
        self._allreduce_rbuf()
        self._scalars(1)
        self._after_init()
        scal = solvers.S_NARR * c
        k, done = 0, 0.0
        if self.use_graph and self.graph is None and self.max_iter >= 2 * self.check:
            self._iteration()                       # eager once: kernel attributes set, NCCL warmed up
            k = 1
            torch.cuda.synchronize()
            before = _lib.launch_count()
            self.graph = solvers._capture(self._iteration, self.check)
            self.launches_per_replay = _lib.launch_count() - before      # kernels of ours inside one replay
            _lib._dll.mgp_add_launch_count(-self.launches_per_replay)    # the capture pass itself launched nothing
        # the done flag of chunk j is read while chunk j + 1 is already enqueued (see solvers.linear_cg)
        flags = torch.zeros(solvers._FLAG_DEPTH, dtype=self.dt).pin_memory()
        events = [torch.cuda.Event() for _ in range(solvers._FLAG_DEPTH)]
        pending, slot = [], 0
        while k < self.max_iter and done == 0.0:
            steps = min(self.check, self.max_iter - k)
            if self.graph is not None and steps == self.check:
                self.graph.replay()
                _lib._dll.mgp_add_launch_count(self.launches_per_replay)
            else:
                for _ in range(steps):
                    self._iteration()
            k += steps
            flags[slot:slot + 1].copy_(self.state[scal + solvers.K_DONE:scal + solvers.K_DONE + 1], non_blocking=True)
            events[slot].record()
            pending.append(slot)
            slot = (slot + 1) % solvers._FLAG_DEPTH
            if len(pending) == solvers._FLAG_DEPTH or self.graph is None:
                s0 = pending.pop(0)
                events[s0].synchronize()
                done = float(flags[s0])
        for s0 in pending:
            if done != 0.0:
                break
            events[s0].synchronize()
            done = float(flags[s0])
        out = torch.empty((n_loc, c), dtype=self.dt, device=self.dev)
        _lib.call("mgp_cg_finalize_" + sfx, ptr(self.x), c_int64(ld), ptr(out), c_int64(c), c_int64(n_loc), c_int32(c),
                  ptr(self.state), stream())
        tail = self.state[scal:scal + 3].tolist()
        info = dict(iterations=int(tail[solvers.K_ITER]), mean_residual=float(tail[solvers.K_MEAN]), converged=(done == 1.0))
        return out, info

    def apply(self, x_loc: torch.Tensor) -> torch.Tensor:
        """(P x) on this rank's rows, outside the iteration: halo all-to-all + the chained local SpMM launches (the NCCL
        transport of ``DistPrecision.matvec``; bare precision operator).  Every rank must call it."""
        op, c = self.op, self.c
        bufs = self.__dict__.get("_apply_bufs")
        if bufs is None:
            z = lambda rows: torch.zeros((rows, self.ld), dtype=self.dt, device=self.dev)
            bufs = self._apply_bufs = (z(op.n_ext), z(op.n_ext), z(op.n_loc))
        src, tmp, out = bufs
        src[:op.n_loc, :c].copy_(x_loc)
        op.matvec(src, out, tmp, c)
        return out[:, :c]

    def solve_polished(self, b_loc: torch.Tensor, factor: float = 8.0):
        """``solve`` plus the one true-residual correction of ``settings.cg_polish`` (solvers._polish) on the partitioned solver:
        fp32 mBCG stops on the recurrence residual, whose drift leaves a TRUE residual of ~2e-4 at cfg-C; evaluate r = b - P x once
        (global norms: two all-reduced sums), solve P d = r to ``factor * tol * |b| / |r|`` with the same kernels, return x + d.
        info gains ``polish = {true_residual_before, iterations}``.  Bare precision operator, fp32, converged solves only."""
        x, info = self.solve(b_loc)
        if self.dt != torch.float32 or self.tol > 1e-4 or not info["converged"] or self.op.coef is not None:
            return x, info
        r = b_loc - self.apply(x)
        sums = torch.stack((r.double().square().sum(0), b_loc.double().square().sum(0)))
        if dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(sums, group=self.group)
        before = float((sums[0] / sums[1].clamp_min(1e-300)).sqrt().mean())
        info["polish"] = {"true_residual_before": before, "iterations": 0}
        if not (before > factor * self.tol) or before != before:
            return x, info
        tol0 = self.tol
        self.tol = min(0.5, factor * tol0 / before)          # the tolerance is written to the device state by the solve's init step
        try:
            d, dinfo = self.solve(r.contiguous())
        finally:
            self.tol = tol0
        info["polish"]["iterations"] = int(dinfo["iterations"])
        return x + d, info


class PeerMemory:
    """Peer-mapped ("symmetric") device allocations of one process group: every rank allocates the same shape and gets the
    device pointers of all ranks' copies (torch.distributed._symmetric_memory: CUDA VMM handles exchanged once at
    rendezvous, NVLink P2P loads / stores afterwards).  Plumbing only -- the data path is the kernels that use the pointers."""

    def __init__(self, device, group=None, world=None):
        self.device = device
        self.world = int(world) if world is not None else dist.get_world_size(group)
        self.handles = []
        if self.world == 1:
            self._sm, self.group = None, group
            return
        import torch.distributed._symmetric_memory as symm_mem
        self._sm = symm_mem
        self.group = group if group is not None else dist.group.WORLD

    def alloc(self, shape, dtype):
        """(local tensor, int64 device tensor with the world's base pointers)."""
        if self.world == 1:      # one rank: its own memory is the whole "symmetric" allocation
            t = torch.zeros(*shape, dtype=dtype, device=self.device)
            self.handles.append(t)
            return t, torch.tensor([t.data_ptr()], dtype=torch.int64, device=self.device)
        t = self._sm.empty(*shape, dtype=dtype, device=self.device)
        hdl = self._sm.rendezvous(t, self.group)
        t.zero_()
        ptrs = torch.tensor(list(hdl.buffer_ptrs), dtype=torch.int64, device=self.device)
        self.handles.append(hdl)
        return t, ptrs

    def sync(self):
        torch.cuda.synchronize(self.device)
        if self.world > 1:
            dist.barrier(group=self.group)


class PeerCG(DistCG):
    """DistCG whose every exchange is a hand-written peer-memory kernel instead of an NCCL call (8 x B200 behind NVSwitch).

    Halo rows are not exchanged at all: the vectors the SpMM reads live in peer-mapped memory and its producer warps cp.async
    the halo rows straight from the owning rank's vector (``peer_x`` of mgp_lap_spmm_wi).  Three generations, ``mode``:

    * ``"cg1"`` (default) -- single-reduction iteration (Chronopoulos-Gear recurrences, csrc/cg.cu::cg_peer_cgstep_kernel):
      nu SpMM launches with r as the source + ONE vector kernel = nu + 1 launches and nu + 1 sync points, every one of
      them published by the LAST block of the producing kernel (no launch latency on the critical path) and waited for
      only where the data is needed (first remote halo row / start of the vector kernel).  Same iterates, masks, stopping
      rules and tridiagonal history as ``solvers.linear_cg``.
    * ``"fused"`` (round 1) -- standard two-reduction CG: barriers inside the SpMM launches (published at kernel start), the
      two all-reduces inside the r / (p, x) update kernels: nu + 2 launches, nu + 2 sync points.
    * ``"unfused"`` -- separate barrier / all-reduce kernels (8 launches), kept as the simplest correct form.
    """

    def __init__(self, op: DistPrecision, ncols: int, dtype=torch.float32, tolerance=1e-6, eps=1e-10,
                 stop_updating_after=1e-10, max_iter=1000, check_interval=16, group=None, use_cuda_graph=True, mode=None,
                 n_tridiag=0, max_tridiag_iter=20):
        super().__init__(op, ncols, dtype, tolerance, eps, stop_updating_after, max_iter, check_interval, group, use_cuda_graph)
        from . import solvers
        dev = self.dev
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        part = op.plan.part
        n_sym = max(b1 - b0 for b0, b1 in zip(part.bounds[:-1], part.bounds[1:]))       # same allocation size on every rank
        self.mem = PeerMemory(dev, group, self.world)
        if mode is None:
            mode = os.environ.get("MGP_PEER_MODE", "cg1")
            if os.environ.get("MGP_PEER_FUSED") == "0":
                mode = "unfused"
        pow2 = self.ld <= 128 and (self.ld & (self.ld - 1)) == 0
        if mode in ("cg1", "fused") and not pow2:
            mode = "unfused"
        self.mode = mode
        self.fused = mode == "fused"
        self.n_tridiag = int(n_tridiag)
        self.n_tridiag_iter = int(min(max_tridiag_iter, part.n)) if n_tridiag else 0
        self.max_hist = max(1, min(max_iter, self.n_tridiag_iter)) if n_tridiag else 0
        self.hist = torch.zeros((self.max_hist, 2, ncols), dtype=dtype, device=dev) if n_tridiag else None
        if n_tridiag and mode != "cg1":
            raise RuntimeError("PeerCG: the tridiagonal history is recorded by the single-reduction mode only")
        self.p, self.p_ptrs = self.mem.alloc((n_sym, self.ld), dtype)
        nstage = op.nstage if mode == "cg1" else op.nu
        if mode != "cg1" and op.coef is not None:
            raise RuntimeError("PeerCG: Scale / Noise wrappers are covered by the single-reduction mode only")
        # intermediate vectors: a buffer may be rewritten once every peer has finished the stage that read it, which the flag
        # waits guarantee two stages later -> three rotating buffers serve any chain length
        self.tmps = [self.mem.alloc((n_sym, self.ld), dtype) for _ in range(min(max(nstage - 1, 0), 3))]
        self.red, self.red_ptrs = self.mem.alloc((2 * self.world * 128,), dtype)
        self.flags, self.flag_ptrs = self.mem.alloc((64,), torch.int32)
        self.epoch = torch.zeros(1, dtype=torch.int32, device=dev)
        self.rbuf_pap = torch.zeros(ncols, dtype=dtype, device=dev)
        self.tmp = None
        scal = solvers.S_NARR * ncols
        self.iter_scalar = self.state[scal + solvers.K_ITER:]
        self.done_scalar = self.state[scal + solvers.K_DONE:]
        if mode in ("cg1", "fused"):
            self.red2, self.red2_ptrs = self.mem.alloc((2 * 2 * self.world * 128,), dtype)
            nflag = nstage + 2
            self.flags2, base = self.mem.alloc((nflag, 64), torch.int32)
            self.flag_tabs = [(base + 64 * 4 * i).contiguous() for i in range(nflag)]
        if mode == "cg1":
            # r is the vector the peers read (source of the first SpMM); p, s, x, w stay local
            self.r, self.r_ptrs = self.mem.alloc((n_sym, self.ld), dtype)
            self.s = torch.zeros((op.n_loc, self.ld), dtype=dtype, device=dev)
            self.gamma_loc = torch.zeros(ncols, dtype=dtype, device=dev)
            self.tickets = torch.zeros(8 * max(nstage, 1), dtype=torch.int32, device=dev)     # one 32-byte slot per SpMM stage
            self.publish_at_end = os.environ.get("MGP_PEER_PUBLISH", "start") == "end"
        self.mem.sync()

    # -- building blocks -----------------------------------------------------------------------------------------------
    def _barrier(self):
        from . import _lib
        from ._lib import c_int32, ptr, stream
        _lib.call("mgp_peer_barrier", ptr(self.flag_ptrs), ptr(self.epoch), c_int32(self.rank), c_int32(self.world), stream())

    def _scalars(self, what, rbuf=None):
        from . import _lib
        from ._lib import c_double, c_float, c_int32, ptr, stream
        fl = c_float if self.dt == torch.float32 else c_double
        rbuf = self.rbuf if rbuf is None else rbuf
        _lib.call("mgp_cg_peer_scalars_" + _lib.suffix(self.dt), ptr(self.state), ptr(rbuf), c_int32(self.c), c_int32(what),
                  fl(self.tol), fl(self.eps), fl(self.stop), c_int32(self.max_iter), c_int32(self.n_tridiag_iter), None, c_int32(0),
                  ptr(self.red_ptrs), ptr(self.flag_ptrs), ptr(self.epoch), c_int32(self.rank), c_int32(self.world), stream())

    def _matvec(self):
        """v <- P p with p^T v partial sums in rbuf_pap; halo rows come from the peers' copies of the source vector."""
        from . import graph
        op, c, n_loc = self.op, self.c, self.op.n_loc
        src, src_ptrs = self.p, self.p_ptrs
        for s in range(op.nu):
            last = s == op.nu - 1
            dst, dst_ptrs = (self.v, None) if last else self.tmps[s]
            if not self.fused:
                self._barrier()                                  # the source vector is complete on every rank
            graph.lap_spmm(op.st, op.a, op.diag, src[:n_loc, :c], shift=op.shift, out=dst[:n_loc, :c],
                           dot_with=self.p[:n_loc, :c] if last else None, dot_out=self.rbuf_pap if last else None,
                           peer_x=src_ptrs,
                           peer_sync=(self.rank, self.flag_tabs[s], self.iter_scalar) if self.fused else None)
            src, src_ptrs = dst, dst_ptrs

    def _matvec_cg1(self):
        """w (= self.v) <- A r for A = P, c P or Noise(c P); the last launch ships (r.w, |r|^2) partials to every rank.  Flag
        tables: [0] "r complete", [i] "output of stage i - 1 complete", [nstage] "dot partials shipped"."""
        from . import _lib, graph
        op, c, n_loc = self.op, self.c, self.op.n_loc
        eps_ = op.stage_epilogues()
        nstage = len(eps_)
        src, src_ptrs = self.r, self.r_ptrs
        # paired-row walk (fp32, whole 64-byte rows) when the operator keeps the paired value stream alive (PartitionedPrecision)
        qrow = op.pair_rows() if (self.dt == torch.float32 and c % 16 == 0 and hasattr(op, "pair_rows")) else None
        for s in range(nstage):
            last = s == nstage - 1
            dst, dst_ptrs = (self.v, None) if last else self.tmps[s % len(self.tmps)]
            ep_coef, ep_add = eps_[s]
            if self.publish_at_end:      # A/B form: flags / partials leave from the END of the producing launch
                ext = _lib.wi_ext(done_flag=self.done_scalar, wait_flags=self.flag_tabs[s],
                                  publish_flags=None if last else self.flag_tabs[s + 1], ticket=self.tickets[8 * s:],
                                  red_ptrs=self.red2_ptrs if last else None, red_flags=self.flag_tabs[nstage] if last else None,
                                  ship_extra=self.gamma_loc if last else None, ship_ncols=c if last else 0,
                                  ep_coef=ep_coef, ep_add=ep_add, pair_rows=qrow)
            else:                        # default: "source complete" published by block 0 of the CONSUMING launch at its start
                ext = _lib.wi_ext(done_flag=self.done_scalar, wait_flags=self.flag_tabs[s], publish_at_start=True,
                                  ep_coef=ep_coef, ep_add=ep_add, pair_rows=qrow)
            graph.lap_spmm(op.st, op.a, op.diag, src[:n_loc, :c], shift=op.shift, out=dst[:n_loc, :c],
                           dot_with=self.r[:n_loc, :c] if (last or ep_add) else None, dot_out=self.rbuf_pap if last else None,
                           peer_x=src_ptrs, peer_ext=(self.rank, self.iter_scalar, ext))
            src, src_ptrs = dst, dst_ptrs

    def _iteration(self):
        from . import _lib
        from ._lib import c_int32, c_int64, ptr, stream
        sfx = _lib.suffix(self.dt)
        n_loc, c, ld = self.op.n_loc, self.c, self.ld
        if self.mode == "cg1":
            nu = self.op.nstage
            self._matvec_cg1()
            _lib.call("mgp_cg_peer_cgstep_" + sfx, ptr(self.x), ptr(self.r), ptr(self.p), ptr(self.s), ptr(self.v), c_int64(ld),
                      c_int64(n_loc), c_int32(c), ptr(self.state), ptr(self.hist), c_int32(self.max_hist), ptr(self.ws),
                      ptr(self.gamma_loc), ptr(None if self.publish_at_end else self.rbuf_pap), ptr(self.red2_ptrs),
                      ptr(self.flag_tabs[nu]), ptr(self.flag_tabs[0] if self.publish_at_end else None),
                      c_int32(self.rank), c_int32(self.world), stream())
            return
        self._matvec()
        if self.fused:
            nu = self.op.nu
            _lib.call("mgp_cg_peer_rupdate_" + sfx, ptr(self.r), ptr(self.v), c_int64(ld), c_int64(n_loc), c_int32(c),
                      ptr(self.state), ptr(self.rbuf_pap), ptr(self.ws), ptr(self.red2_ptrs), ptr(self.flag_tabs[nu]),
                      ptr(self.flag_tabs[nu + 1]), c_int32(self.rank), c_int32(self.world), stream())
            _lib.call("mgp_cg_peer_pxupdate_" + sfx, ptr(self.x), ptr(self.p), ptr(self.r), c_int64(ld), c_int64(n_loc),
                      c_int32(c), ptr(self.state), None, c_int32(0), ptr(self.ws), ptr(self.red2_ptrs),
                      ptr(self.flag_tabs[nu + 1]), c_int32(self.rank), c_int32(self.world), stream())
            return
        self._scalars(3, self.rbuf_pap)                          # all-reduce(p^T A p) -> state
        _lib.call("mgp_cg_rupdate_" + sfx, ptr(self.r), ptr(self.v), c_int64(ld), c_int64(n_loc), c_int32(c), ptr(self.state),
                  None, c_int32(0), ptr(self.rbuf), ptr(self.ws), stream())
        self._scalars(2)                                         # all-reduce(|r|^2) -> alpha, beta, flags
        _lib.call("mgp_cg_pxupdate_" + sfx, ptr(self.x), ptr(self.p), ptr(self.r), c_int64(ld), c_int64(n_loc), c_int32(c),
                  ptr(self.state), stream())

    def _allreduce_rbuf(self):
        pass                                                     # fused into _scalars

    def _after_init(self):
        """cg1: the initial residual is complete on this rank -> local |r_0|^2 for the first shipment, s = 0, publish r."""
        if self.mode != "cg1":
            return
        from . import _lib
        from ._lib import c_int32, ptr, stream
        import ctypes
        self.gamma_loc.copy_(self.rbuf)                          # cg_dist_init left this rank's |r_0|^2 column sums there
        self.s.zero_()
        if self.publish_at_end:
            _lib.call("mgp_peer_publish", ptr(self.flag_tabs[0]), ctypes.c_uint32(1), c_int32(self.rank), c_int32(self.world), stream())

    def solve(self, b_loc: torch.Tensor):
        if self.mode in ("cg1", "fused"):
            # the iteration-keyed flags restart from zero every solve: nobody may still be publishing into them (barrier),
            # and nobody may publish before everyone has zeroed (barrier)
            self._barrier()
            self.flags2.zero_()
            if self.hist is not None:
                self.hist.zero_()
            self._barrier()
        return super().solve(b_loc)

    def tridiagonals(self, info):
        """Lanczos tridiagonals of the first ``n_tridiag`` columns from the recorded CG coefficients (identical on every rank)."""
        from . import solvers
        k_done = info["iterations"]
        rows = k_done - 1 if info["converged"] else k_done
        rows = max(1, min(rows, self.n_tridiag_iter))
        return solvers._tridiag_from_hist(self.hist.cpu(), rows, self.n_tridiag, self.dt).to(self.dev)


def dist_lanczos_tridiag(dop: "DistPrecision", max_iter: int, init_vec_loc: torch.Tensor, tol: float = 1e-5, group=None):
    """Row-partitioned Lanczos with full re-orthogonalisation (``solvers.lanczos_tridiag`` over several GPUs; the reference's
    ``GraphLaplacianOperator.diagonalization`` path, graph_laplacian_operator.py:132-135, SURVEY.md 8e "CG / Lanczos").

    ``dop``: a ``DistPrecision`` (with ``nu = 1`` and ``shift = 0`` it is the Laplacian itself).  Every rank holds its rows of
    the Lanczos vectors; per step: halo exchange + local fused SpMV, local ``Q^T r`` partial dots (``mgp_lanczos_dots``),
    ONE all-reduce of ``[j]`` floats, local ``r -= Q c`` + norm partial (``mgp_lanczos_axpy``), one all-reduce of the norm;
    check passes as in the single-GPU driver.  Returns (q_loc [steps, n_loc], t [steps, steps]) -- ``t`` identical on every
    rank, so ``eigh(t)`` and the Ritz vectors ``q_loc^T V`` need no further communication."""
    from . import _lib
    from ._lib import c_int32, c_int64, ptr, stream
    n_loc, n_ext = dop.n_loc, dop.n_ext
    dt, dev = init_vec_loc.dtype, init_vec_loc.device
    sfx = _lib.suffix(dt)
    n_glob = dop.plan.part.n
    num_iter = min(int(max_iter), n_glob)
    q = torch.zeros((num_iter + 1, n_loc), dtype=dt, device=dev)
    src = torch.zeros((n_ext, 1), dtype=dt, device=dev)              # [own rows | halo rows] of the current Lanczos vector
    tmp = torch.zeros((n_ext, 1), dtype=dt, device=dev)
    r = init_vec_loc.reshape(n_loc).to(dt).clone()
    c = torch.zeros(num_iter + 1, dtype=dt, device=dev)
    nrm2 = torch.zeros(1, dtype=dt, device=dev)
    beta = torch.zeros(1, dtype=dt, device=dev)
    ws = torch.zeros(_lib.query("mgp_lanczos_ws_bytes", c_int64(n_loc), c_int32(num_iter + 1)), dtype=torch.uint8, device=dev)
    alphas = torch.zeros(num_iter, dtype=dt, device=dev)
    betas = torch.zeros(num_iter, dtype=dt, device=dev)
    world = dist.get_world_size(group)

    def allreduce(t):
        if world > 1:
            dist.all_reduce(t, group=group)

    def dots(j):
        _lib.call("mgp_lanczos_dots_" + sfx, ptr(q), c_int64(n_loc), c_int32(j), ptr(r), c_int64(n_loc), ptr(c), ptr(ws), stream())
        allreduce(c[:max(j, 1)])

    def axpy(j):
        _lib.call("mgp_lanczos_axpy_" + sfx, ptr(q), c_int64(n_loc), c_int32(j), ptr(r), c_int64(n_loc), ptr(c), ptr(nrm2), ptr(ws), stream())
        allreduce(nrm2)

    def normalize(dst, beta_out):
        _lib.call("mgp_lanczos_normalize_" + sfx, ptr(r), c_int64(n_loc), ptr(nrm2), ptr(dst), ptr(beta_out), stream())

    c.zero_()
    axpy(0)                                                            # j = 0: norm only
    normalize(q[0], None)
    steps = 0
    for k in range(num_iter):
        src[:n_loc, 0].copy_(q[k])
        dop.matvec(src, r.unsqueeze(-1), tmp, 1)                       # r <- A q_k on this rank's rows (halo exchange inside)
        j = k + 1
        dots(j)
        axpy(j)                                                        # removes alpha_k q_k, beta_{k-1} q_{k-1} and every other component
        alphas[k:k + 1].copy_(c[k:k + 1])
        steps = k + 1
        if k + 1 >= num_iter:
            break
        ok = False
        nr2 = 0.0
        for _ in range(10):
            dots(j)
            worst, nr2 = torch.stack((c[:j].abs().max(), nrm2[0])).tolist()
            if nr2 <= 0.0 or worst / (nr2 ** 0.5) <= tol:
                ok = True
                break
            axpy(j)
        normalize(q[k + 1], beta)
        betas[k:k + 1].copy_(beta)
        if nr2 ** 0.5 <= 1e-6 or not ok:
            break
    t = torch.diag(alphas[:steps])
    if steps > 1:
        off = betas[:steps - 1]
        t = t + torch.diag(off, 1) + torch.diag(off, -1)
    return q[:steps], t


class DistBackend:
    """Routes ``solvers.linear_cg`` of the package's native operators -- PrecisionMaternOperator, bare or inside the reference's
    Scale / Noise wrappers (riemann_gp.py:32-39) -- to the row-partitioned peer-memory CG of this process group, so that
    ``manifold_informed_train`` / ``inv_quad_logdet`` / ``solve`` / ``_average_variance`` run UNCHANGED on every rank with their
    CG + SLQ solves (the dominant cost: utils/train_model.py:55,67-68) sharded over the GPUs:

        from manifold_gp_b200 import distributed, solvers
        solvers.set_distributed_backend(distributed.DistBackend())      # after init_process_group, same call on every rank

    Every rank passes the same full right-hand side; rows are partitioned in the structure's Morton order; the solution is
    all-gathered (one collective per solve, outside the iteration) and the CG coefficients / Lanczos tridiagonals are identical
    on every rank, so the SLQ log-det, the surrogate backward and the optimiser step run redundantly and stay in lock-step.
    Graph structure and value build are replicated per rank (each rank holds the whole graph: 8 GB at N = 10M, k = 32).
    Partition-dependent objects (halo plan, local views, peer-mapped vectors, captured CUDA graphs) are cached per graph
    structure and reused across bandwidths: a training step only copies the new values into the solver's buffers."""

    def __init__(self, group=None, mode="cg1", min_rows=100_000):
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.mode = mode
        self.min_rows = int(min_rows)
        self._cache = {}
        self.solves = 0

    @staticmethod
    def _decompose(op):
        """(precision operator, coef tensor or None, noise tensor or None) or None when the operator is not covered."""
        from .operators.noise_wrapper_operator import NoiseWrapperOperator
        from .operators.precision_matern_operator import PrecisionMaternOperator
        from .operators.scale_wrapper_operator import ScaleWrapperOperator
        noise = coef = None
        cur = op
        if isinstance(cur, NoiseWrapperOperator):
            noise, cur = cur.noise, cur.operator
        if isinstance(cur, ScaleWrapperOperator):
            sc = cur.scale.detach().reshape(-1)[:1]
            coef, cur = (sc.reciprocal() if cur.inverse_scale else sc), cur.operator
        if not isinstance(cur, PrecisionMaternOperator) or not cur._native():
            return None
        if cur.laplacian.normalization != "symmetric":
            return None                       # the D^{+-1/2} factors of the random-walk form are not partitioned yet
        return cur, coef, noise

    def supports(self, op, rhs) -> bool:
        if not rhs.is_cuda or rhs.shape[0] < self.min_rows:
            return False
        dec = self._decompose(op)
        if dec is None:
            return False
        st = dec[0].laplacian.structure
        t = st.build_tiles()
        return t is not None and "wptr" in t and rhs.shape[1] <= 128

    def sharded_self_search(self, knn, k):
        """kNN of the cloud against itself: this rank searches its block of query rows, the lists are all-gathered
        (nearest_neighbors.py:35-37 at multi-GPU scale; identical results on every rank)."""
        x = knn._db
        n = x.shape[0]
        part = RowPartition(n, self.world, align=64)
        lo, hi = part.range(self.rank)
        t0 = time.perf_counter()
        d_loc, i_loc = knn.search(x[lo:hi], k)           # a row slice: not the self-search branch again
        torch.cuda.synchronize()
        self.last_knn_local_s = time.perf_counter() - t0
        sizes = [part.range(r)[1] - part.range(r)[0] for r in range(self.world)]
        d_all = [torch.empty((sz, k), dtype=d_loc.dtype, device=x.device) for sz in sizes]
        i_all = [torch.empty((sz, k), dtype=i_loc.dtype, device=x.device) for sz in sizes]
        dist.all_gather(d_all, d_loc.contiguous(), group=self.group)
        dist.all_gather(i_all, i_loc.contiguous(), group=self.group)
        return torch.cat(d_all), torch.cat(i_all)

    def lanczos_eigenpairs(self, lap, max_iter, init_vec=None, tol=1e-5):
        """``solvers.diagonalization(method="lanczos")`` of a symmetric-normalised GraphLaplacianOperator over the process group:
        (evals ascending [j], evecs [n, j] in the caller's row order, identical on every rank)."""
        from . import solvers
        gst = lap.structure
        n = lap.shape[0]
        with torch.no_grad():
            _, _, diag, a = lap._values()
            diag, a = diag.detach(), a.detach()
        part = RowPartition(n, self.world, align=gst.TILE_ROWS)
        dop = DistPrecision(gst, diag, a, torch.zeros(1, dtype=a.dtype, device=a.device), 1, part, self.rank, self.group)
        lo, hi = part.range(self.rank)
        if init_vec is None:
            init_vec = torch.randn(n, dtype=a.dtype, device=a.device)
            if self.world > 1:
                dist.broadcast(init_vec, 0, group=self.group)
        q_loc, t = dist_lanczos_tridiag(dop, max_iter, gst.to_internal(init_vec.reshape(n, 1))[lo:hi, 0], tol, self.group)
        evals, v = solvers.lanczos_tridiag_to_diag(t)
        ev_loc = q_loc.T @ v
        if self.world > 1:
            sizes = [part.range(r)[1] - part.range(r)[0] for r in range(self.world)]
            parts = [torch.empty((sz, ev_loc.shape[1]), dtype=ev_loc.dtype, device=ev_loc.device) for sz in sizes]
            dist.all_gather(parts, ev_loc.contiguous(), group=self.group)
            ev_loc = torch.cat(parts)
        return evals, gst.to_external(ev_loc), t

    def cg_chunk(self, op, rhs, n_tridiag, tolerance, eps, stop_updating_after, max_iter, n_tridiag_iter):
        """Same contract as ``solvers._cg_chunk``: (solution [n, C] in the caller's row order, hist (cpu) or None, info)."""
        from . import solvers
        prec, coef, noise = self._decompose(op)
        lap = prec.laplacian
        gst = lap.structure
        c_in = rhs.shape[1]
        # the fused vector kernels want a power-of-two row length: zero columns are free in a 64-byte-row pass and have
        # residual exactly 0, so tolerance * c / cpad is the published mean-over-columns rule on the original block
        cpad = 4
        while cpad < c_in:
            cpad *= 2
        if cpad != c_in:
            wide = torch.zeros((rhs.shape[0], cpad), dtype=rhs.dtype, device=rhs.device)
            wide[:, :c_in] = rhs
            sol, hist, info = self.cg_chunk(op, wide, n_tridiag, tolerance * c_in / cpad, eps, stop_updating_after, max_iter, n_tridiag_iter)
            info["mean_residual"] = info["mean_residual"] * cpad / c_in
            info["residual_norm"] = info["residual_norm"][:c_in]
            return sol[:, :c_in], (hist[:, :, :c_in] if hist is not None else None), info
        n, c = rhs.shape
        dt = rhs.dtype
        with torch.no_grad():
            _, _, diag, a = lap._values()
            diag, a = diag.detach(), a.detach()
            shift = prec._shift_const(a.dtype)
        ent = self._cache.get(id(gst))
        if ent is None or ent["gst"] is not gst:
            part = RowPartition(n, self.world, align=gst.TILE_ROWS)
            ent = {"gst": gst, "part": part, "ops": {}, "solvers": {}}
            self._cache = {id(gst): ent}                      # one graph at a time: the buffers are per-rank gigabytes
        okey = (dt, prec.nu, coef is not None, noise is not None)
        dop = ent["ops"].get(okey)
        if dop is None:
            dop = DistPrecision(gst, diag, a, shift, prec.nu, ent["part"], self.rank, self.group, coef=coef, noise=noise)
            ent["ops"][okey] = dop
        else:
            dop.update_values(diag, a, shift, coef, noise)
        skey = okey + (c, int(n_tridiag), int(n_tridiag_iter))
        cg = ent["solvers"].get(skey)
        if cg is None:
            cg = PeerCG(dop, c, dt, tolerance=tolerance, eps=eps, stop_updating_after=stop_updating_after, max_iter=max_iter,
                        check_interval=int(solvers.settings.cg_check_interval.value()), group=self.group, mode=self.mode,
                        n_tridiag=n_tridiag, max_tridiag_iter=n_tridiag_iter if n_tridiag else 20)
            ent["solvers"][skey] = cg
        cg.tol, cg.eps, cg.stop, cg.max_iter = float(tolerance), float(eps), float(stop_updating_after), int(max_iter)
        part = ent["part"]
        lo, hi = part.range(self.rank)
        b_loc = gst.to_internal(rhs)[lo:hi].contiguous()
        x_loc, info = cg.solve(b_loc)
        sizes = [part.range(r)[1] - part.range(r)[0] for r in range(self.world)]
        if self.world > 1:
            parts = [torch.empty((sz, c), dtype=dt, device=rhs.device) for sz in sizes]
            dist.all_gather(parts, x_loc.contiguous(), group=self.group)
            sol = torch.cat(parts)
        else:
            sol = x_loc
        sol = gst.to_external(sol)
        from .solvers import CGInfo, S_RESID
        cinfo = CGInfo(iterations=info["iterations"], mean_residual=info["mean_residual"], converged=info["converged"],
                       residual_norm=cg.state[S_RESID * c:(S_RESID + 1) * c].clone(), distributed=self.world)
        self.solves += 1
        return sol, (cg.hist.cpu() if n_tridiag else None), cinfo


def _collectives_note(cg, transport, nu):
    if transport != "peer":
        return {"halo_all_to_all": nu, "all_reduce": 2}
    mode = getattr(cg, "mode", "unfused")
    halo = "read from the owners' vectors over NVLink inside the SpMM"
    if mode == "cg1":
        at_end = bool(getattr(cg, "publish_at_end", False))
        return {"launches": nu + 1, "sync_points": nu + 1,
                "all_reduce": "1 (Chronopoulos-Gear single-reduction CG): (r.w, |r|^2) partials shipped by "
                              + ("the last block of the last SpMM launch" if at_end else "block 0 of the one vector kernel at its start")
                              + ", added in rank order by every block of that kernel",
                "barrier": f"{nu} 'source vector complete' flags, published "
                           + ("by the last block of the PRODUCING kernel" if at_end else "by block 0 of the CONSUMING SpMM launch at its start")
                           + ", waited for at the consumer's first remote halo row", "halo": halo}
    if mode == "fused":
        return {"launches": nu + 2, "sync_points": nu + 2,
                "barrier": "inside the SpMM launches (flags over NVLink, published at kernel start, waited on at the first remote halo row)",
                "all_reduce": "2, inside the r / (p, x) update kernels (partials shipped to every peer)", "halo": halo}
    return {"peer_barrier": nu, "peer_allreduce_fused_with_scalars": 2, "halo": halo}


def dist_cg(op: DistPrecision, b_loc: torch.Tensor, tolerance=1e-6, eps=1e-10, stop_updating_after=1e-10, max_iter=1000,
            check_interval=16, group=None, use_cuda_graph=False):
    """One-shot convenience wrapper around :class:`DistCG` (no graph caching)."""
    return DistCG(op, b_loc.shape[1], b_loc.dtype, tolerance, eps, stop_updating_after, max_iter, check_interval, group,
                  use_cuda_graph).solve(b_loc)


def sharded_knn(x: torch.Tensor, k: int, part_queries: RowPartition, rank: int, group=None):
    """kNN with the database replicated and the queries sharded; returns the full (dist2, idx) on every rank."""
    from .utils import NearestNeighbors
    lo, hi = part_queries.range(rank)
    knn = NearestNeighbors(x)
    d_loc, i_loc = knn.search(x[lo:hi].contiguous(), k)
    world = part_queries.world
    sizes = [part_queries.range(r)[1] - part_queries.range(r)[0] for r in range(world)]
    d_all = [torch.empty((s, k), dtype=d_loc.dtype, device=x.device) for s in sizes]
    i_all = [torch.empty((s, k), dtype=i_loc.dtype, device=x.device) for s in sizes]
    dist.all_gather(d_all, d_loc.contiguous(), group=group)
    dist.all_gather(i_all, i_loc.contiguous(), group=group)
    return torch.cat(d_all), torch.cat(i_all), knn


# ---------------------------------------------------------------------------------------------------------------------------
# bench entry (called by bench.py under torchrun)
# ---------------------------------------------------------------------------------------------------------------------------
def bench_main(args, CFG, clock_sampler=None):
    import manifold_gp_b200 as mgp
    from manifold_gp_b200 import _lib, graph
    from manifold_gp_b200.utils import synthetic
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ.get("LOCAL_RANK", rank))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    warnings.simplefilter("ignore")
    n, k, c = args.n, CFG["k"], CFG["rhs"]
    x = synthetic.torus(n, seed=CFG["seed"], device=dev)
    # ---- graph ----------------------------------------------------------------------------------------------------------
    # "partitioned" (default): a rank searches, symmetrises and builds structure + values for ITS rows only (PartitionedGraph);
    # "replicated": the round-1 form -- queries sharded, symmetrise / structure / values on every rank for the whole graph.
    build = os.environ.get("MGP_DIST_BUILD", "partitioned")
    pg = None
    if build == "partitioned":
        torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
        pg = PartitionedGraph(x, k)
        torch.cuda.synchronize(); dist.barrier(); t_knn = time.perf_counter() - t0
        eps = float(pg.gather(pg.kth_dist2.unsqueeze(1)).squeeze(1).sqrt().median())
        op = PartitionedPrecision(pg, eps, CFG["nu"], CFG["kappa"], CFG["self_loops"])
        part = pg.part
        lo, hi = part.range(rank)
        nnz_t = torch.tensor([pg.st.nnz], dtype=torch.int64, device=dev)
        dist.all_reduce(nnz_t)
        m = int(nnz_t.item()) // 2
        g = torch.Generator(device=dev).manual_seed(CFG["rhs_seed"])
        B = torch.randn(n, c, device=dev, generator=g)           # identical on every rank (same seed)
        b_loc = pg.to_local(B).contiguous()
    else:
        # ---- graph: queries sharded, everything after replicated ------------------------------------------------------------
        qpart = RowPartition(n, world, align=64)
        torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
        d2, nbr, knn = sharded_knn(x, k, qpart, rank)
        torch.cuda.synchronize(); dist.barrier(); t_knn = time.perf_counter() - t0
        eps = float(d2[:, k - 1].sqrt().median())
        # symmetrise on every rank (deterministic, identical results)
        from ._lib import c_int32, c_int64, c_size_t, ptr, stream
        cap = n * (k - 1)
        eidx = torch.empty((2, cap), dtype=torch.int64, device=dev)
        ev = torch.empty(cap, dtype=torch.float32, device=dev)
        m_out = torch.zeros(1, dtype=torch.int64, device=dev)
        wsb = _lib.workspace(_lib.query("mgp_graph_symmetrize_ws_bytes", c_int64(n), c_int32(k)), dev)
        _lib.call("mgp_graph_symmetrize_f32", ptr(d2.contiguous()), ptr(nbr.contiguous()), c_int64(n), c_int32(k), c_int32(1),
                  ptr(eidx), ptr(ev), c_int64(cap), ptr(m_out), ptr(wsb), c_size_t(wsb.numel()), stream())
        m = int(m_out.item())
        idx, val = eidx[:, :m].contiguous(), ev[:m].contiguous()
        del eidx, ev, wsb
        graph.attach_permutation(idx, graph.morton_permutation(x))
        lap = mgp.GraphLaplacianOperator(val, idx, n, torch.tensor([[eps]], device=dev), CFG["normalization"], CFG["self_loops"])
        prec = mgp.PrecisionMaternOperator(lap, CFG["nu"], torch.tensor([[CFG["kappa"]]], device=dev))
        gst = lap.structure
        _, _, diag, a = lap._values()
        part = RowPartition(n, world, align=gst.TILE_ROWS)
        op = DistPrecision(gst, diag, a, prec._shift(), CFG["nu"], part, rank)
        lo, hi = part.range(rank)
        g = torch.Generator(device=dev).manual_seed(CFG["rhs_seed"])
        B = torch.randn(n, c, device=dev, generator=g)           # identical on every rank (same seed)
        b_loc = gst.to_internal(B)[lo:hi].contiguous()


    transport = os.environ.get("MGP_DIST_TRANSPORT", "peer")
    if transport == "peer":
        cg = PeerCG(op, c, torch.float32, tolerance=CFG["tol"], max_iter=CFG["max_iter"])
    else:
        cg = DistCG(op, c, torch.float32, tolerance=CFG["tol"], max_iter=CFG["max_iter"])

    def solve():
        # like the single-GPU arm: the timed solve includes the one true-residual correction (settings.cg_polish)
        return cg.solve_polished(b_loc)

    for _ in range(args.warmup):
        xs, info = solve()
    torch.cuda.synchronize(); dist.barrier()
    _lib.reset_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    clk = None
    if clock_sampler is not None and rank == 0:            # SM clocks / throttle reasons during the timed region (rank 0's GPU)
        try:
            clk = clock_sampler(local)
            clk.__enter__()
        except Exception:
            clk = None
    ev0.record()
    for _ in range(args.steps):
        xs, info = solve()
    ev1.record()
    torch.cuda.synchronize(); dist.barrier()
    clocks = None
    if clk is not None:
        try:
            clk.__exit__(None, None, None)
            clocks = clk.summary()
        except Exception:
            clocks = None
    ms = torch.tensor([ev0.elapsed_time(ev1) / args.steps], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    launches = _lib.launch_count()
    # check against the global residual and against the SINGLE-GPU solve of the same system (every rank holds the replicated
    # operator; rank 0 runs solvers.linear_cg on it after the timed region)
    if pg is not None:
        sol = pg.gather(xs)
        res = b_loc - cg.apply(xs)                               # true residual from the partitioned operator itself
        sums = torch.stack((res.double().square().sum(0), b_loc.double().square().sum(0)))
        dist.all_reduce(sums)
        true_rel = float((sums[0] / sums[1]).sqrt().mean())
        prec = None
        if rank == 0:                                            # the replicated operator exists on rank 0 only, for the check below
            idx, val = mgp.NearestNeighbors(x).graph(k)
            lap = mgp.GraphLaplacianOperator(val, idx, n, torch.tensor([[eps]], device=dev), CFG["normalization"], CFG["self_loops"])
            prec = mgp.PrecisionMaternOperator(lap, CFG["nu"], torch.tensor([[CFG["kappa"]]], device=dev))
    else:
        x_all = [torch.empty((part.range(r)[1] - part.range(r)[0], c), device=dev) for r in range(world)]
        dist.all_gather(x_all, xs.contiguous())
        sol = gst.to_external(torch.cat(x_all))
        true_rel = float(((prec.matmul(sol) - B).double().norm(dim=0) / B.double().norm(dim=0)).mean())
    vs_single = None
    if rank == 0:
        from . import settings, solvers
        with settings.cg_polish(False):
            ref, rinfo = solvers.linear_cg(prec, B, tolerance=CFG["tol"], max_iter=CFG["max_iter"], return_info=True)
        err = ((sol - ref).double().norm(dim=0) / ref.double().norm(dim=0))
        vs_single = {"single_gpu_iterations": int(rinfo["iterations"]), "solution_rel_diff_max": float(err.max()),
                     "solution_rel_diff_mean": float(err.mean()), "within_1e-4": bool(float(err.max()) <= 1e-4)}
        del ref
    dist.barrier()
    # e2e: host buffers in, host result out (each rank moves its own block)
    Bh = b_loc.cpu().pin_memory(); Xh = torch.empty_like(Bh).pin_memory()
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    for _ in range(max(1, min(args.steps, 2))):
        bl = Bh.to(dev, non_blocking=True)
        xs2, _ = cg.solve_polished(bl)
        Xh.copy_(xs2, non_blocking=True)
        torch.cuda.synchronize()
    dist.barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / max(1, min(args.steps, 2))
    if rank == 0:
        polish = dict(info.get("polish") or {})
        iters = int(info["iterations"]) + int(polish.get("iterations", 0))
        unit = "CG iterations/s (N=1M, k=32, nu=2, 16 RHS per iteration)"
        out = {"metric": "precision_cg_iterations_per_s", "value": round(iters / (float(ms) * 1e-3), 1), "unit": unit, "n_gpus": world,
               "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(float(ms), 3), "solve_ms": round(float(ms), 3),
               "higher_is_better": True,
               "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
               "config": {"workload": CFG["workload"], "n": n, "k": k, "edges_M": m, "nnz": 2 * m, "nu": CFG["nu"], "kappa": CFG["kappa"],
                          "eps": round(eps, 6), "rhs": c, "tol": CFG["tol"], "normalization": CFG["normalization"],
                          "self_loops": CFG["self_loops"], "partition": "contiguous row blocks of the Morton order",
                          "graph_build": ("partitioned: every rank searches, symmetrises (one all_to_all_v) and builds structure + values "
                                          "(two halo gathers) for its rows only" if pg is not None else
                                          "replicated: queries sharded, symmetrise / structure / values on every rank"),
                          "entries_rank0": int(op.st.nnz if pg is not None else 0) or None,
                          "halo_rows_rank0": int(op.plan.halo_ids.numel()), "rows_rank0": int(op.n_loc),
                          "l2": "inputs larger than L2 at 1-2 GPUs; at 8 GPUs a rank's share (~50 MB) is L2 resident (strong scaling)"},
               "cg_iterations": int(info["iterations"]), "cg_polish_iterations": int(polish.get("iterations", 0)),
               "cg_converged": bool(info["converged"]) and true_rel <= 10 * CFG["tol"], "cg_converged_recurrence": bool(info["converged"]),
               "cg_true_relative_residual": true_rel, "cg_true_relative_residual_unpolished": polish.get("true_residual_before"), "parity_vs_single_gpu": vs_single, "knn_build_s": round(t_knn, 4),
               "e2e": {"value": round(iters / (e2e_ms * 1e-3), 1), "unit": unit, "solve_ms": round(e2e_ms, 3),
                       "h2d_bytes_per_step": int(Bh.numel() * 4 * world), "d2h_bytes_per_step": int(Xh.numel() * 4 * world)},
               "gpu_launches": int(launches),
               "clocks": clocks,
               "transport": transport,
               "peer_mode": getattr(cg, "mode", None),
               "collectives_per_iteration": _collectives_note(cg, transport, CFG["nu"])}
        print(json.dumps(out))
    dist.barrier()
    return None


# ---------------------------------------------------------------------------------------------------------------------------
# Partitioned graph construction (SURVEY.md 8e rows "symmetrise + coalesce" and "value build"): a rank builds, holds and
# updates ONLY its rows -- no rank ever sees the whole edge list.
# ---------------------------------------------------------------------------------------------------------------------------
def partitioned_symmetrize(rows: torch.Tensor, cols: torch.Tensor, vals: torch.Tensor, part: RowPartition, rank: int, group=None):
    """Row-partitioned form of ``NearestNeighbors.graph``'s symmetrise + mean-coalesce (nearest_neighbors.py:45-51).

    In: the directed kNN edges of the rows this rank owns (``rows`` in [lo, hi), ``cols`` anywhere, ``vals`` squared distances,
    self column already dropped).  The reference maps an edge to (min, max) and averages duplicates; row i of the symmetric
    operator then lists every j with i -> j or j -> i.  Here ONE ``all_to_all_v`` sends each edge i -> j to the owner of j as
    (j, i, d): afterwards a rank holds both directions of every edge incident to its rows and coalesces them locally -- the
    mean over the duplicates of {i, j} is taken from exactly the directed edges the reference averages.  A surviving
    ``j == i`` (duplicate points) is listed twice, as the reference's two scatter passes count it.
    Returns (row, col, val) of the rank's rows, sorted by (row, col), columns in the GLOBAL numbering.  Device-agnostic."""
    world = part.world
    dev = rows.device
    rows, cols = rows.to(torch.int64), cols.to(torch.int64)
    owner = part.owner(cols).clamp_(max=world - 1)
    notdiag = cols != rows                          # i -> i is its own reverse: it already sits in the out-list
    o_nd = owner[notdiag]
    order = torch.argsort(o_nd, stable=True)
    send_ids = torch.stack((cols[notdiag][order], rows[notdiag][order]), dim=1).contiguous()      # (their row, their column)
    send_vals = vals[notdiag][order].contiguous()
    send_counts = torch.bincount(o_nd, minlength=world)
    recv_counts = torch.empty_like(send_counts)
    dist.all_to_all_single(recv_counts, send_counts, group=group)
    sc, rc = send_counts.tolist(), recv_counts.tolist()
    recv_ids = torch.empty((int(sum(rc)), 2), dtype=torch.int64, device=dev)
    recv_vals = torch.empty(int(sum(rc)), dtype=vals.dtype, device=dev)
    dist.all_to_all_single(recv_ids, send_ids, rc, sc, group=group)
    dist.all_to_all_single(recv_vals, send_vals, rc, sc, group=group)
    r = torch.cat((rows, recv_ids[:, 0]))
    c = torch.cat((cols, recv_ids[:, 1]))
    v = torch.cat((vals, recv_vals))
    key = r * part.n + c
    ukey, inv, cnt = torch.unique(key, return_inverse=True, return_counts=True)
    acc = torch.zeros(ukey.numel(), dtype=v.dtype, device=dev).scatter_add_(0, inv, v)
    val = acc / cnt.to(v.dtype)
    row, col = ukey // part.n, ukey % part.n
    rep = 1 + (row == col).to(torch.int64)          # diagonal entries count twice
    if bool((rep > 1).any()):
        row, col, val = row.repeat_interleave(rep), col.repeat_interleave(rep), val.repeat_interleave(rep)
    return row, col, val


def _all_gather_blocks(block: torch.Tensor, part: RowPartition, group=None) -> torch.Tensor:
    """Concatenation of every rank's row block (blocks of a RowPartition differ in length: the last one is shorter).  Padded to a
    common length so that any backend's equal-size all_gather serves (gloo has no uneven form)."""
    world = part.world
    if world == 1:
        return block
    sizes = [part.range(r)[1] - part.range(r)[0] for r in range(world)]
    mx = max(sizes)
    padded = torch.zeros((mx,) + tuple(block.shape[1:]), dtype=block.dtype, device=block.device)
    padded[:block.shape[0]] = block
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    return torch.cat([p[:sz] for p, sz in zip(parts, sizes)])


def partitioned_sort_rows(row, col, val, part: RowPartition, rank: int, perm_block: torch.Tensor, tile_rows: int = 128, group=None):
    """Inside every ``tile_rows``-row tile of this rank's block, order the rows by their number of entries (descending) -- what
    GraphStructure does for the replicated structure: the 8 row slots of a warp block walk in lock-step, so they want rows of
    equal length (measured on 2 GPUs at cfg-C: 360 ms without against 343 ms for the replicated build).  The renumbering is local
    to a rank's block, but the other ranks name these rows as halo columns: one halo gather of the new ids translates them, and
    the row order of the whole partition is re-assembled with one all-gather.
    In: the rank's entries (global ``row`` in [lo, hi), global ``col``, ``val``) and ``perm_block`` = original index of each of
    its rows.  Returns (row, col, val) renumbered and sorted by (row, col), and the refined ``perm`` [n] of ALL rows.
    Device-agnostic (collectives: the HaloPlan handshake, one exchange, one all-gather)."""
    lo, hi = part.range(rank)
    n_loc, dev = hi - lo, row.device
    world = part.world
    deg = torch.bincount(row - lo, minlength=n_loc)
    ar = torch.arange(n_loc, device=dev, dtype=torch.int64)
    key = ((ar // tile_rows) << 20) | ((1 << 20) - 1 - deg.clamp_max((1 << 20) - 1))
    order = torch.argsort(key, stable=True)                      # new local position -> old local row
    newpos = torch.empty_like(order)
    newpos[order] = ar
    plan0 = HaloPlan(part, rank, col, group=group)
    ids = torch.zeros((n_loc + int(plan0.halo_ids.numel()), 1), dtype=torch.int64, device=dev)
    ids[:n_loc, 0] = lo + newpos
    plan0.exchange(ids)                                          # new global ids of the halo rows, from their owners
    col = ids[plan0.to_local(col), 0]
    row = lo + newpos[row - lo]
    o = torch.argsort(row * part.n + col, stable=True)
    block = perm_block[order].contiguous()                       # my block of the refined row order
    perm = _all_gather_blocks(block, part, group)
    return row[o], col[o], val[o], perm


class PartitionedGraph:
    """The kNN graph of a point cloud, ROW-PARTITIONED AT CONSTRUCTION (one process per GPU): the database is replicated (as in
    ``sharded_knn``: the exhaustive search needs all points), everything after the search is partitioned --

    * rows are contiguous, tile-aligned blocks of the Morton order of the points (computed redundantly, deterministic);
    * a rank searches the neighbours of ITS rows only, exchanges the reverse edges (``partitioned_symmetrize``: one
      all_to_all_v of (row, col, dist^2) triples) and builds the CSR / tile / stream structure of its rows
      (``GraphStructure.from_rows``; columns in the extended numbering [own rows | halo rows] of its ``HaloPlan``);
    * ``values`` runs the three passes of the value build on its rows with two halo gathers (degrees, then normalised
      degrees) between them (``mgp_lap_values_pass``).

    Per rank: nnz / world entries instead of the whole structure (8 GB at N = 10M, k = 32 when replicated)."""

    def __init__(self, x: torch.Tensor, k: int, group=None):
        from . import graph
        from .utils import NearestNeighbors
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        n = int(x.shape[0])
        self.n, self.k = n, int(k)
        self.perm = graph.morton_permutation(x)                       # new position -> original index
        self.inv = torch.empty_like(self.perm)
        self.inv[self.perm] = torch.arange(n, device=x.device)
        self.part = RowPartition(n, self.world, align=graph.GraphStructure.TILE_ROWS)
        lo, hi = self.part.range(self.rank)
        self.lo, self.hi = lo, hi
        knn = NearestNeighbors(x)
        d2, nbr = knn.search(x.index_select(0, self.perm[lo:hi]).contiguous(), k)
        self.knn_info = knn.last_search
        self.kth_dist2 = d2[:, k - 1].contiguous()                   # bandwidth heuristics (median k-th neighbour distance)
        rows = torch.arange(lo, hi, device=x.device, dtype=torch.int64).repeat_interleave(k - 1)
        cols = self.inv[nbr[:, 1:].reshape(-1).to(torch.int64)]
        row, col, val = partitioned_symmetrize(rows, cols, d2[:, 1:].reshape(-1), self.part, self.rank, group)
        n_loc = hi - lo
        row, col, val = self._sort_rows_inside_tiles(row, col, val, n_loc)
        self.plan = HaloPlan(self.part, self.rank, col, group=group)
        rowptr = torch.searchsorted(row, torch.arange(lo, hi + 1, device=x.device, dtype=torch.int64))
        self.st = graph.GraphStructure.from_rows(rowptr, self.plan.to_local(col), n_loc, n_loc + int(self.plan.halo_ids.numel()))
        self.entry_d2 = val.contiguous()                             # per entry, in the order the structure's ``eid`` indexes
        self.n_loc, self.n_ext = n_loc, self.st.n_cols
        t = self.st.tiles
        if t is not None and "hcol" in t:
            # peer-memory variant of the halo lists: (owner rank << 26) | row inside the owner's block
            loc = t["hcol"].to(torch.int64)
            gid = torch.where(loc < n_loc, loc + lo, self.plan.halo_ids[(loc - n_loc).clamp(min=0, max=max(self.plan.halo_ids.numel() - 1, 0))]
                              if self.plan.halo_ids.numel() else loc + lo)
            own = self.part.owner(gid).clamp_(max=self.world - 1)
            starts = torch.tensor(self.part.bounds[:-1], dtype=torch.int64, device=x.device)
            t["hcol_peer"] = ((own << 26) | (gid - starts[own])).to(torch.int32).contiguous()

    def _sort_rows_inside_tiles(self, row, col, val, n_loc):
        """See ``partitioned_sort_rows``; also refreshes ``perm`` / ``inv`` with the refined row order of the whole partition."""
        from . import graph
        row, col, val, self.perm = partitioned_sort_rows(row, col, val, self.part, self.rank, self.perm[self.lo:self.hi],
                                                         graph.GraphStructure.TILE_ROWS, self.group)
        self.inv = torch.empty_like(self.perm)
        self.inv[self.perm] = torch.arange(self.n, device=row.device)
        return row, col, val

    def values(self, eps, self_loops: bool = True, dtype=torch.float32):
        """(deg_unnorm [n_ext], deg [n_ext], diag [n_loc], a [nnz_loc]) for bandwidth ``eps`` -- graph_laplacian_operator.py:52-106
        on this rank's rows; the halo parts of the two degree vectors come from their owners (two halo gathers)."""
        from . import _lib, graph
        from ._lib import c_int32, c_int64, ptr, stream
        st = self.st
        d2 = st.d2csr(self.entry_d2.to(dtype))
        dev = st.device
        eps_t = graph._device_scalar(eps, dtype, dev)
        dt = torch.zeros(self.n_ext, dtype=dtype, device=dev)
        dg = torch.zeros(self.n_ext, dtype=dtype, device=dev)
        diag = torch.empty(self.n_loc, dtype=dtype, device=dev)
        a = torch.zeros(st.nnz + 8, dtype=dtype, device=dev)[:st.nnz]
        sfx = _lib.suffix(dtype)
        for p in (1, 2, 3):
            _lib.call("mgp_lap_values_pass_" + sfx, c_int32(p), ptr(st.rowptr), ptr(st.col), ptr(d2), c_int64(self.n_loc), ptr(eps_t),
                      c_int32(1 if self_loops else 0), ptr(dt), ptr(dg), ptr(diag), ptr(a), stream())
            if p == 1:
                self.plan.exchange(dt.view(-1, 1))
            elif p == 2:
                self.plan.exchange(dg.view(-1, 1))
        return dt, dg, diag, a

    # rows of a replicated [n, C] array that this rank owns, in the structure's order / back
    def to_local(self, v: torch.Tensor) -> torch.Tensor:
        return v.index_select(0, self.perm[self.lo:self.hi])

    def gather(self, v_loc: torch.Tensor) -> torch.Tensor:
        """[n, C] in the caller's point order from every rank's block (one all-gather)."""
        return _all_gather_blocks(v_loc.contiguous(), self.part, self.group).index_select(0, self.inv)


class PartitionedPrecision(DistPrecision):
    """``DistPrecision`` on a ``PartitionedGraph``: same surface for the partitioned solvers (``DistCG`` / ``PeerCG`` /
    ``dist_lanczos_tridiag``), but structure and values are built from this rank's rows only; ``update`` re-runs the
    partitioned value build into the same buffers (captured CUDA graphs and peer-memory registrations survive)."""

    def __init__(self, pg: PartitionedGraph, eps, nu: int, kappa, self_loops: bool = True, dtype=torch.float32, coef=None, noise=None):
        self.pg, self.plan, self.st, self.gst = pg, pg.plan, pg.st, None
        self.lo, self.hi, self.nu = pg.lo, pg.hi, int(nu)
        self.n_loc, self.n_ext = pg.n_loc, pg.n_ext
        self.self_loops = bool(self_loops)
        dev = pg.st.device
        self.diag = torch.empty(self.n_loc, dtype=dtype, device=dev)
        self.a = torch.zeros(pg.st.nnz + 8, dtype=dtype, device=dev)[:pg.st.nnz]
        self.shift = torch.empty(1, dtype=dtype, device=dev)
        self.coef = torch.ones(1, dtype=dtype, device=dev) if (coef is not None or noise is not None) else None
        self.ncoef = torch.zeros(1, dtype=dtype, device=dev) if noise is not None else None
        self.has_noise = noise is not None
        t = pg.st.tiles
        if t is not None and "wptr" in t:
            self._aw = torch.zeros(t["nnzw"] + 64, dtype=dtype, device=dev)
        else:
            self._aw = None
        # paired-row streams (fp32): the value stream lives in ONE buffer across bandwidths, like the single-row one
        self._aq = None
        if self._aw is not None and dtype == torch.float32 and os.environ.get("MGP_DIST_PAIR", "1") != "0":
            tq = pg.st.pair_tiles()
            if tq is not None:
                self._aq = torch.zeros(tq["qsrc"].numel(), dtype=dtype, device=dev)
        if self._aq is None:
            pg.st._pair_tried = True        # no paired walk on this structure (auto dispatch of the NCCL-transport matvec included)
        self.update(eps, kappa, coef, noise)

    def pair_rows(self):
        """Row table of the paired-row streams when this operator keeps their value stream alive, else None."""
        return self.st.tiles["qrow"] if self._aq is not None else None

    def update(self, eps, kappa, coef=None, noise=None):
        """New bandwidth / lengthscale / scales: same memory, new contents."""
        from . import graph
        with torch.no_grad():
            _, _, diag, a = self.pg.values(eps, self.self_loops, self.a.dtype)
            self.diag.copy_(diag)
            self.a.copy_(a)
            kap = graph._device_scalar(kappa, self.a.dtype, self.a.device)
            self.shift.copy_((2.0 * self.nu) / kap.square())
            if self._aw is not None:
                self.st.__dict__["_aw_persistent"] = None
                self._aw.copy_(self.st.wi_values(self.a))
                self.st._aw_persistent = self._aw
            if self._aq is not None:
                self.st.__dict__["_aq_persistent"] = None
                self._aq.copy_(self.st.pair_values(self.a))
                self.st._aq_persistent = self._aq
            if self.coef is not None:
                c = torch.ones(1, dtype=self.coef.dtype, device=self.coef.device) if coef is None else \
                    graph._device_scalar(coef, self.coef.dtype, self.coef.device)
                self.coef.copy_(c)
                if self.ncoef is not None:
                    self.ncoef.copy_(-graph._device_scalar(noise, self.coef.dtype, self.coef.device) * c)

    def update_values(self, *a, **k):  # the replicated-values entry of the base class does not apply
        raise RuntimeError("PartitionedPrecision: use update(eps, kappa, ...) -- values are built from this rank's rows")
