"""Row-partitioned multi-GPU execution of the Matern-precision CG solve.

The reference has no distributed code; this is new design.  One process per GPU (torchrun), ``torch.distributed`` (NCCL
over NVLink/NVSwitch; gloo in the CPU tests) for the plumbing:

* rows are partitioned into contiguous, 128-aligned blocks of the space-filling-curve order, so a rank's neighbours are
  almost all its own rows and the halo is the thin boundary of a 2-D patch;
* every matvec exchanges ONLY halo rows of the vector (``all_to_all_single`` with precomputed send lists; the receive
  buffer is the halo region of the local vector, ordered by owner so each peer's rows land contiguously);
* every CG inner product is an all-reduce of ``C`` floats (p^T A p comes out of the SpMM epilogue, r^T r out of the fused
  update kernel; ``mgp_cg_dist_scalars`` finishes alpha/beta/flags on every rank identically);
* graph construction is replicated in round 1 (kNN queries are sharded and all-gathered; the O(nnz) structure / value
  builds run on every rank): only the solve is partitioned.

The local operator is a *view* of the global tile-compacted structure (no index rebasing: row pointers keep global
positions so the TMA copies stay 16-byte aligned) with the halo column ids translated to local vector rows.
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
        return self._gst.wi_values(a)    # built from the GLOBAL rowptr / wptr: stream positions are global


class DistPrecision:
    """(2nu/kappa^2 + L_sym)^nu on the rows of this rank.  ``values`` = (diag[n], a[nnz]) of the GLOBAL structure."""

    def __init__(self, gst, diag, a, shift, nu: int, part: RowPartition, rank: int, group=None):
        lo, hi = part.range(rank)
        p0, p1 = int(gst.rowptr[lo]), int(gst.rowptr[hi])
        self.plan = HaloPlan(part, rank, gst.col[p0:p1], group=group)
        self.st = LocalStructure(gst, self.plan)
        self.diag = diag[lo:hi]
        self.a = a
        self.shift = shift
        self.nu = nu
        self.n_loc = hi - lo
        self.n_ext = self.n_loc + int(self.plan.halo_ids.numel())

    def matvec(self, p, out, tmp, ncols, dot_out=None):
        """out[:n_loc, :ncols] <- P p  (p, tmp: [n_ext, ld] with halo rows; out: [>= n_loc, ld])."""
        from . import graph
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


class PeerMemory:
    """Peer-mapped ("symmetric") device allocations of one process group: every rank allocates the same shape and gets the
    device pointers of all ranks' copies (torch.distributed._symmetric_memory: CUDA VMM handles exchanged once at
    rendezvous, NVLink P2P loads / stores afterwards).  Plumbing only -- the data path is the kernels that use the pointers."""

    def __init__(self, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        self._sm = symm_mem
        self.device = device
        self.group = group if group is not None else dist.group.WORLD
        self.handles = []

    def alloc(self, shape, dtype):
        """(local tensor, int64 device tensor with the world's base pointers)."""
        t = self._sm.empty(*shape, dtype=dtype, device=self.device)
        hdl = self._sm.rendezvous(t, self.group)
        t.zero_()
        ptrs = torch.tensor(list(hdl.buffer_ptrs), dtype=torch.int64, device=self.device)
        self.handles.append(hdl)
        return t, ptrs

    def sync(self):
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)


class PeerCG(DistCG):
    """DistCG whose every exchange is a hand-written peer-memory kernel instead of an NCCL call (8 x B200 behind NVSwitch):

    * halo rows are not exchanged at all: ``p`` and the intermediate vectors live in peer-mapped memory and the SpMM's
      producer warps cp.async the halo rows straight from the owning rank's vector (``peer_x`` of mgp_lap_spmm_wi);
      producer and consumer launches are separated by ``mgp_peer_barrier`` (one tiny kernel, epoch flags over NVLink);
    * the two inner products of an iteration are all-reduced inside the scalar kernels (``mgp_cg_peer_scalars``).

    An iteration is 8 kernel launches and no library call; NCCL's launch + protocol latency (~37 us per 16-float
    all-reduce, ~2 x that per halo all-to-all on this box) was ~2/3 of the 8-GPU iteration time."""

    def __init__(self, op: DistPrecision, ncols: int, dtype=torch.float32, tolerance=1e-6, eps=1e-10,
                 stop_updating_after=1e-10, max_iter=1000, check_interval=16, group=None, use_cuda_graph=True):
        super().__init__(op, ncols, dtype, tolerance, eps, stop_updating_after, max_iter, check_interval, group, use_cuda_graph)
        from . import solvers
        dev = self.dev
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        part = op.plan.part
        n_sym = max(b1 - b0 for b0, b1 in zip(part.bounds[:-1], part.bounds[1:]))       # same allocation size on every rank
        self.mem = PeerMemory(dev, group)
        self.p, self.p_ptrs = self.mem.alloc((n_sym, self.ld), dtype)
        self.tmps = [self.mem.alloc((n_sym, self.ld), dtype) for _ in range(max(op.nu - 1, 0))]
        self.red, self.red_ptrs = self.mem.alloc((2 * self.world * 128,), dtype)
        self.flags, self.flag_ptrs = self.mem.alloc((64,), torch.int32)
        self.epoch = torch.zeros(1, dtype=torch.int32, device=dev)
        self.rbuf_pap = torch.zeros(ncols, dtype=dtype, device=dev)
        self.tmp = None
        # fused iteration (default): the cross-GPU barriers ride inside the SpMM launches and the two all-reduces inside the
        # r / (p, x) update kernels, all keyed by the iteration counter in ``state`` -> 2 nu + 2... = nu + 2 launches.
        import os
        self.fused = os.environ.get("MGP_PEER_FUSED", "1") != "0" and self.ld <= 128 and (self.ld & (self.ld - 1)) == 0
        if self.fused:
            self.red2, self.red2_ptrs = self.mem.alloc((2 * 2 * self.world * 128,), dtype)
            nflag = op.nu + 2                                        # one flag array per SpMM stage + p^T A p + |r|^2
            self.flags2, base = self.mem.alloc((nflag, 64), torch.int32)
            self.flag_tabs = [(base + 64 * 4 * i).contiguous() for i in range(nflag)]
            self.iter_scalar = self.state[solvers.S_NARR * ncols + solvers.K_ITER:]
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
                  fl(self.tol), fl(self.eps), fl(self.stop), c_int32(self.max_iter), c_int32(0), None, c_int32(0),
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

    def _iteration(self):
        from . import _lib
        from ._lib import c_int32, c_int64, ptr, stream
        sfx = _lib.suffix(self.dt)
        n_loc, c, ld = self.op.n_loc, self.c, self.ld
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

    def solve(self, b_loc: torch.Tensor):
        if self.fused:
            # the iteration-keyed flags restart from zero every solve: nobody may still be publishing into them (barrier),
            # and nobody may publish before everyone has zeroed (barrier)
            self._barrier()
            self.flags2.zero_()
            self._barrier()
        return super().solve(b_loc)


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
        return cg.solve(b_loc)

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
    # check against the global residual
    x_all = [torch.empty((part.range(r)[1] - part.range(r)[0], c), device=dev) for r in range(world)]
    dist.all_gather(x_all, xs.contiguous())
    sol = gst.to_external(torch.cat(x_all))
    true_rel = float(((prec.matmul(sol) - B).double().norm(dim=0) / B.double().norm(dim=0)).mean())
    # e2e: host buffers in, host result out (each rank moves its own block)
    Bh = b_loc.cpu().pin_memory(); Xh = torch.empty_like(Bh).pin_memory()
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    for _ in range(max(1, min(args.steps, 2))):
        bl = Bh.to(dev, non_blocking=True)
        xs2, _ = cg.solve(bl)
        Xh.copy_(xs2, non_blocking=True)
        torch.cuda.synchronize()
    dist.barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / max(1, min(args.steps, 2))
    if rank == 0:
        out = {"metric": "precision_cg_solve_time", "value": round(float(ms), 3), "unit": "ms", "n_gpus": world,
               "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(float(ms), 3), "higher_is_better": False,
               "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
               "config": {"workload": CFG["workload"], "n": n, "k": k, "edges_M": m, "nu": CFG["nu"], "kappa": CFG["kappa"],
                          "eps": round(eps, 6), "rhs": c, "tol": CFG["tol"], "partition": "contiguous row blocks of the Morton order",
                          "halo_rows_rank0": int(op.plan.halo_ids.numel()), "rows_rank0": int(op.n_loc),
                          "l2": "inputs larger than L2 at 1-2 GPUs; at 8 GPUs a rank's share (~50 MB) is L2 resident (strong scaling)"},
               "cg_iterations": info["iterations"], "cg_converged": bool(info["converged"]),
               "cg_true_relative_residual": true_rel, "knn_build_s": round(t_knn, 4),
               "e2e": {"value": round(e2e_ms, 3), "unit": "ms", "h2d_bytes_per_step": int(Bh.numel() * 4 * world),
                       "d2h_bytes_per_step": int(Xh.numel() * 4 * world)},
               "gpu_launches": int(launches),
               "clocks": clocks,
               "transport": transport,
               "collectives_per_iteration": (
                   ({"launches": CFG["nu"] + 2, "barrier": "inside the SpMM launches (flags over NVLink, waited on at the first remote halo row)",
                     "all_reduce": "2, inside the r / (p, x) update kernels (partials shipped to every peer)",
                     "halo": "read from the owners' vectors over NVLink inside the SpMM"} if getattr(cg, "fused", False) else
                    {"peer_barrier": CFG["nu"], "peer_allreduce_fused_with_scalars": 2,
                     "halo": "read from the owners' vectors over NVLink inside the SpMM"})
                   if transport == "peer" else {"halo_all_to_all": CFG["nu"], "all_reduce": 2})}
        print(json.dumps(out))
    dist.barrier()
    return None
