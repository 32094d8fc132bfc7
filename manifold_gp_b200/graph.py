"""Device-side graph structure and thin Python wrappers over the C ABI (libmgp_b200.so).

``GraphStructure`` is the hyper-parameter independent part of the Laplacian: the row-major directed structure
(CSR over both directions of every edge of the reference's upper-triangular COO, nearest_neighbors.py:48-51) built
once per graph and cached on the ``idx`` tensor.  Per-bandwidth values are produced by :func:`lap_values`.
"""
from __future__ import annotations

import os

import torch

from . import _lib
from ._lib import c_int32, c_int64, c_size_t, ptr, stream


def _device_scalar(v, dtype, device) -> torch.Tensor:
    """0-d/1-element device tensor of ``dtype`` holding ``v`` (tensor -> no host sync; python number -> H2D of 1 value)."""
    if torch.is_tensor(v):
        return v.detach().reshape(-1)[:1].to(device=device, dtype=dtype).contiguous()
    return torch.tensor([float(v)], dtype=dtype, device=device)


def _spread21(v: torch.Tensor) -> torch.Tensor:
    """Spread the low 21 bits of an int64 tensor so that two zero bits separate consecutive bits (3-D Morton)."""
    v = (v | (v << 32)) & 0x1F00000000FFFF
    v = (v | (v << 16)) & 0x1F0000FF0000FF
    v = (v | (v << 8)) & 0x100F00F00F00F00F
    v = (v | (v << 4)) & 0x10C30C30C30C30C3
    v = (v | (v << 2)) & 0x1249249249249249
    return v


def morton_permutation(x: torch.Tensor) -> torch.Tensor:
    """Space-filling-curve order of the points (new position -> old index).  Legal because every operator on the path
    is permutation-equivariant; it makes the SpMM's gathers of X rows cache-local (neighbours in space become
    neighbours in memory) and is what keeps a row-partitioned halo small.  For d > 3 the curve runs over the three
    leading principal directions.  Setup only (torch ops): once per graph, hyper-parameter independent."""
    xs = x.detach().to(torch.float32)
    if xs.shape[1] > 3:
        xc = xs - xs.mean(0, keepdim=True)
        _, _, v = torch.pca_lowrank(xc, q=3, center=False)
        xs = xc @ v[:, :3]
    d = xs.shape[1]
    lo, hi = xs.amin(0), xs.amax(0)
    scale = (2 ** 21 - 1) / (hi - lo).clamp_min(1e-30)
    q = ((xs - lo) * scale).to(torch.int64).clamp_(0, 2 ** 21 - 1)
    key = torch.zeros(xs.shape[0], dtype=torch.int64, device=xs.device)
    for j in range(d):
        key |= _spread21(q[:, j]) << j
    return torch.argsort(key, stable=True)


_PERM_ATTR = "_mgp_b200_perm"


def attach_permutation(idx: torch.Tensor, perm: torch.Tensor) -> None:
    """Remember a locality-improving row order on the edge-index tensor (picked up by ``structure_for``)."""
    try:
        setattr(idx, _PERM_ATTR, perm)
    except Exception:  # pragma: no cover
        pass


# ---- graph persistence ----------------------------------------------------------------------------------
# The kNN graph is the most expensive setup step (O(N^2 d)) and is hyper-parameter independent, but the reference rebuilds
# it in every run (riemann_kernel.py:40-42 builds it in __init__, nothing is written to disk).  One file holds what a
# later run needs to skip the search: the coalesced upper-triangular edge list, its mean squared distances, the Morton row
# order the operators use, and a fingerprint of the point cloud it was built from.
GRAPH_FILE_FORMAT = "mgp_b200.knn_graph"
GRAPH_FILE_VERSION = 1


def points_fingerprint(x: torch.Tensor) -> str:
    """SHA-256 over shape, dtype and the raw bytes of up to 2 x 1 MiB of the point cloud (head and tail)."""
    import hashlib
    xc = x.detach().contiguous().cpu()
    raw = xc.view(torch.uint8).reshape(-1)
    h = hashlib.sha256()
    h.update(repr((tuple(xc.shape), str(xc.dtype))).encode())
    lim = 1 << 20
    h.update(raw[:lim].numpy().tobytes())
    h.update(raw[-lim:].numpy().tobytes())
    return h.hexdigest()


def save_knn_graph(path: str, edge_index: torch.Tensor, edge_value: torch.Tensor, n: int, k: int, x: torch.Tensor = None) -> None:
    """Write the graph ``NearestNeighbors.graph`` returned (plus its attached row order) to ``path``."""
    idx = edge_index.detach().cpu()
    perm = getattr(edge_index, _PERM_ATTR, None)
    payload = {
        "format": GRAPH_FILE_FORMAT, "version": GRAPH_FILE_VERSION, "n": int(n), "k": int(k),
        "edge_index": idx.to(torch.int32) if n < 2 ** 31 else idx,        # int32 on disk halves the file; int64 in memory
        "edge_value": edge_value.detach().cpu(),
        "perm": None if perm is None else perm.detach().cpu().to(torch.int32 if n < 2 ** 31 else torch.int64),
        "points_fingerprint": None if x is None else points_fingerprint(x),
    }
    torch.save(payload, path)


def load_knn_graph(path: str, device, x: torch.Tensor = None, k: int = None):
    """(edge_index [2,M] int64, edge_value [M]) on ``device`` with the row order re-attached.  Raises if the file was
    built from a different point cloud (``x`` given) or a different neighbour count (``k`` given)."""
    payload = torch.load(path, map_location="cpu", weights_only=True)
    if payload.get("format") != GRAPH_FILE_FORMAT or payload.get("version") != GRAPH_FILE_VERSION:
        raise ValueError(f"{path}: not a {GRAPH_FILE_FORMAT} v{GRAPH_FILE_VERSION} file")
    if k is not None and int(payload["k"]) != int(k):
        raise ValueError(f"{path}: built with k={payload['k']}, requested k={k}")
    if x is not None:
        if int(payload["n"]) != int(x.shape[0]):
            raise ValueError(f"{path}: built for {payload['n']} points, got {x.shape[0]}")
        fp = payload.get("points_fingerprint")
        if fp is not None and fp != points_fingerprint(x):
            raise ValueError(f"{path}: point cloud fingerprint mismatch (the graph was built from different data)")
    idx = payload["edge_index"].to(device=device, dtype=torch.int64)
    val = payload["edge_value"].to(device)
    if payload["perm"] is not None:
        attach_permutation(idx, payload["perm"].to(device=device, dtype=torch.int64))
    return idx, val


def padded_streams(rowptr: torch.Tensor, lcol16: torch.Tensor, n: int, R: int):
    """Padded entry streams for the pipelined kernel (csrc/lap_spmm_pipe.cu): every row holds a multiple of 4 entries, every
    tile of ``R`` rows starts at a multiple of 8 (the last row of a tile absorbs the tile's slack); padding entries carry
    the row's own tile-local index (their value is 0 in the padded value array).  Pure index plumbing (any device)."""
    dev = rowptr.device
    ntiles = (n + R - 1) // R
    rp = rowptr.to(torch.int64)
    rowlen = rp[1:] - rp[:-1]
    nnz = int(rp[-1])
    plen = (rowlen + 3) & ~3
    rid = torch.arange(n, device=dev, dtype=torch.int64)
    tsum = torch.zeros(ntiles, dtype=torch.int64, device=dev)
    tsum.index_add_(0, rid // R, plen)
    slack = ((tsum + 7) & ~7) - tsum
    last_row = torch.clamp(torch.arange(1, ntiles + 1, device=dev, dtype=torch.int64) * R, max=n) - 1
    plen[last_row] += slack
    prowptr = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    torch.cumsum(plen, 0, out=prowptr[1:])
    nnzp = int(prowptr[-1])
    if nnzp >= 2 ** 31 - 16:
        return None
    plcol = torch.zeros(nnzp + 8, dtype=torch.int16, device=dev)
    plcol[:nnzp] = torch.repeat_interleave((rid % R).to(torch.int16), plen)
    rows = torch.repeat_interleave(rid, rowlen)
    dst = prowptr[:-1][rows] + (torch.arange(nnz, device=dev, dtype=torch.int64) - rp[:-1][rows])
    plcol[dst] = lcol16
    ptile = prowptr[torch.arange(0, n + R, R, device=dev).clamp_max(n)]
    return dict(prowptr=prowptr.to(torch.int32).contiguous(), plcol=plcol, nnzp=nnzp, pnzmax=int((ptile[1:] - ptile[:-1]).max()))


def wi_streams(rowptr: torch.Tensor, lcol16: torch.Tensor, n: int, R: int = 128):
    """Warp-interleaved entry streams of the v5 SpMM kernel (csrc/lap_spmm_wi.cu).  A tile of R = 128 rows is walked by 16
    warps of 8 row slots x 4 lanes; warp block b = 16 * tile + w owns stream positions [wptr[b], wptr[b+1]) and position
    wptr[b] + 32 t + lane holds nonzero 4t + (lane & 3) of row 128 tile + 8 w + (lane >> 2) -- exactly the order in which
    the lanes consume them.  Every block is padded to a whole number of 32-entry steps (the longest row of the block,
    rounded up to 4 entries); padding entries have value 0 and a tile-local index of the parity their position expects
    (the row's neighbour r ^ 1), so they cause no extra bank conflicts.  Pure index plumbing (any device)."""
    assert R == 128
    dev = rowptr.device
    ntiles = (n + R - 1) // R
    nblk = ntiles * 16
    rp = rowptr.to(torch.int64)
    rowlen = torch.zeros(ntiles * R, dtype=torch.int64, device=dev)
    rowlen[:n] = rp[1:] - rp[:-1]
    steps = ((rowlen + 3) >> 2).view(nblk, 8).amax(1)                         # 32-entry steps per warp block
    wptr = torch.zeros(nblk + 1, dtype=torch.int64, device=dev)
    torch.cumsum(steps * 32, 0, out=wptr[1:])
    nnzw = int(wptr[-1])
    if nnzw >= 2 ** 31 - 64:
        return None
    # padding index per position: neighbour row of the slot's row (valid and of opposite parity), 0 for rows beyond n
    blk = torch.repeat_interleave(torch.arange(nblk, device=dev, dtype=torch.int64), steps * 32)
    pos = torch.arange(nnzw, device=dev, dtype=torch.int64)
    lane = (pos - wptr[blk]) & 31
    rloc = (blk & 15) * 8 + (lane >> 2)
    row = (blk >> 4) * R + rloc
    nb = rloc ^ 1
    pad = torch.where((blk >> 4) * R + nb < n, nb, rloc)
    pad = torch.where(row < n, pad, torch.zeros_like(pad))
    wcol = torch.zeros(nnzw + 64, dtype=torch.int16, device=dev)
    wcol[:nnzw] = pad.to(torch.int16)
    del blk, pos, lane, rloc, row, nb, pad
    nnz = int(rp[-1])
    rows = torch.repeat_interleave(torch.arange(n, device=dev, dtype=torch.int64), rowlen[:n])
    e = torch.arange(nnz, device=dev, dtype=torch.int64) - rp[:-1][rows]
    b = (rows // R) * 16 + (rows % R) // 8
    dst = wptr[b] + (e >> 2) * 32 + (rows & 7) * 4 + (e & 3)
    wcol[dst] = lcol16
    tile_tot = wptr[16::16] - wptr[:-1:16]
    # the kernel bulk-copies metadata in chunks of 32 tiles: pad so every chunk copy (516 ints) stays in bounds
    wpad = torch.full((512 * ((ntiles + 31) // 32) + 4,), nnzw, dtype=torch.int32, device=dev)
    wpad[:nblk + 1] = wptr.to(torch.int32)
    return dict(wptr=wpad, wcol=wcol, nnzw=nnzw, wnzmax=int(tile_tot.max()))


def pair_matching(rowptr: torch.Tensor, lcol16: torch.Tensor, n: int, R: int = 128, chunk_tiles: int = 2048):
    """Which rows of a tile share a slot of the paired-row walk: a greedy maximum-weight matching of every tile's rows, weight =
    number of columns two rows have in common (the X-row loads the pair saves).  Per tile the 128 x 128 weight matrix is one
    small product of the tile's 0/1 pattern with its transpose; the matching is the locally-dominant-edge form of the greedy
    algorithm (every round each free row points at its heaviest free partner, mutual choices are matched; ties are broken by
    a strict order on the pairs, so the heaviest remaining edge is always mutual and <= 64 rounds finish a tile).  On the
    cfg-C torus: 0.594 union entries per nonzero against 0.661 for Morton-adjacent rows (measured on a 60k proxy).
    Returns ``pos`` [ntiles * R] int64: position of every (real or virtual) row in its tile's pair order -- rows at positions
    2p, 2p + 1 are a pair.  Setup only (torch ops, any device), once per graph."""
    dev = rowptr.device
    ntiles = (n + R - 1) // R
    rp = rowptr.to(torch.int64)
    rowlen = rp[1:] - rp[:-1]
    rows = torch.repeat_interleave(torch.arange(n, device=dev, dtype=torch.int64), rowlen)
    lc = lcol16.to(torch.int64) & 0xFFFF
    L = int(lc.max()) + 1 if lc.numel() else R
    tile_ptr = rp[torch.arange(0, ntiles + 1, device=dev, dtype=torch.int64).mul(R).clamp_max(n)]
    wdt = torch.float16 if dev.type == "cuda" else torch.float32
    ar = torch.arange(R, device=dev, dtype=torch.int64)
    tb = torch.minimum(ar[:, None], ar[None, :]) * R + torch.maximum(ar[:, None], ar[None, :])      # strict order on unordered pairs
    pos = torch.empty(ntiles * R, dtype=torch.int64, device=dev)
    # bound the dense 0/1 pattern of a chunk (T x R x L) to ~1 GiB: graphs without locality have wide tiles
    chunk_tiles = max(1, min(int(chunk_tiles), (1 << 30) // (R * L * (2 if wdt == torch.float16 else 4))))
    for c0 in range(0, ntiles, chunk_tiles):
        c1 = min(ntiles, c0 + chunk_tiles)
        T = c1 - c0
        e0, e1 = int(tile_ptr[c0]), int(tile_ptr[c1])
        pat = torch.zeros(T, R, L, dtype=wdt, device=dev)
        pat[rows[e0:e1] // R - c0, rows[e0:e1] % R, lc[e0:e1]] = 1
        W = torch.bmm(pat, pat.transpose(1, 2)).to(torch.float32).to(torch.int64) * (R * R) + tb     # [T, R, R], symmetric
        del pat
        W.diagonal(dim1=1, dim2=2).fill_(-1)
        alive = torch.ones(T, R, dtype=torch.bool, device=dev)
        partner = torch.full((T, R), -1, dtype=torch.int64, device=dev)
        for _ in range(R // 2):
            best = W.masked_fill(~alive[:, None, :], -1).argmax(2)
            mutual = alive & alive.gather(1, best) & (best.gather(1, best) == ar)
            partner = torch.where(mutual, best, partner)
            alive &= ~mutual
            if not bool(alive.any()):
                break
        del W
        # pair order: by the smaller row of the pair; the smaller row takes the even position
        lo = torch.minimum(ar.expand(T, R), partner)
        order = torch.argsort(lo * (2 * R) + ar, dim=1, stable=True)                 # rows of a tile sorted by (pair key, row)
        pr = torch.empty(T, R, dtype=torch.int64, device=dev)
        pr.scatter_(1, order, ar.expand(T, R).contiguous())
        pos[c0 * R:c1 * R] = pr.reshape(-1)
    return pos


def pair_streams(rowptr: torch.Tensor, lcol16: torch.Tensor, n: int, spatial_pos: torch.Tensor = None, R: int = 128):
    """Paired-row entry streams of the v6 SpMM walk (csrc/lap_spmm_wi.cu, PAIR = true).  Two spatially adjacent rows of a tile
    share most of their columns (kNN graph in Morton order: the union of a pair's lists is ~0.66 of their sum), so a slot of
    8 lanes walks the UNION list of a row pair: one 64-byte X-row load from shared memory feeds both rows (32 FMAs instead
    of 16 per load -- the walk is bound by shared-memory wavefronts).  Layout, per tile of R = 128 rows:

    * the 64 pairs are the rows at positions 2p, 2p + 1 of ``spatial_pos`` (position of every row in its tile's pair order:
      ``pair_matching``, or the Morton order of the rows; None = rows 2p, 2p + 1),
      ordered by union length (descending) so the 4 pairs of a warp block run in lock-step; ``qrow[128 tile + 2 pi + h]`` is the
      tile-local row that half h of pair pi outputs (uint8);
    * warp block b = 16 tile + (pi >> 2), slot = pi & 3; stream position qptr[b] + 32 t + 8 slot + l holds the union entry that
      lane l of the slot consumes in step t: its tile-local column in ``qcol`` and TWO source positions into the CSR value
      array in ``qsrc[2 pos], qsrc[2 pos + 1]`` = (value of the lane's OWN output row, value of the other row), -1 = zero.
      Lanes 0-3 of a slot output row A (h = 0), lanes 4-7 row B, so the first exchange of the final reduction needs no select;
    * bank conflicts: lane l reads 16-byte chunk (c + l) & 3 in load c, lanes l and l + 4 of a slot read the same chunk of
      different X rows -- conflict-free when the two columns have opposite parity, so lanes 0-3 are filled with even
      columns and lanes 4-7 with odd ones (the surplus of one parity spills into the other side's free positions);
    * padding entries (value sources -1) carry a valid own-row index of the parity their lane expects.

    Pure index plumbing (any device); duplicates of a column inside one row (a diagonal COO entry is listed twice) are kept
    as separate union entries."""
    assert R == 128
    dev = rowptr.device
    ntiles = (n + R - 1) // R
    nblk = ntiles * 16
    NT = ntiles * R
    rp = rowptr.to(torch.int64)
    nnz = int(rp[-1])
    rowlen = rp[1:] - rp[:-1]
    ar = lambda m: torch.arange(m, device=dev, dtype=torch.int64)
    # spatial position of every (real or virtual) row inside its tile -> first-cut pair id and half
    q0 = ar(NT) % R
    if spatial_pos is not None and spatial_pos.numel() == NT:          # a full pair order (pair_matching): virtual rows included
        q0 = spatial_pos.to(torch.int64) % R
    elif spatial_pos is not None:
        q0[:n] = spatial_pos.to(torch.int64) % R
        if NT > n:                                                   # virtual rows take the positions the real rows left free
            used = torch.zeros(R, dtype=torch.bool, device=dev)
            used[q0[(ntiles - 1) * R:n]] = True
            q0[n:] = (~used).nonzero().reshape(-1)
    gp_row = (ar(NT) // R) * (R // 2) + (q0 >> 1)                    # global pair id of every row (before the length sort)
    half_row = q0 & 1
    rows = torch.repeat_interleave(ar(n), rowlen)
    lc = lcol16.to(torch.int64) & 0xFFFF
    # occurrence rank of a column inside its row (0 except for repeated entries)
    k1 = rows * 65536 + lc
    o1 = torch.argsort(k1, stable=True)
    k1s = k1[o1]
    first = torch.ones(nnz, dtype=torch.bool, device=dev)
    first[1:] = k1s[1:] != k1s[:-1]
    start = torch.cummax(torch.where(first, ar(nnz), torch.zeros((), dtype=torch.int64, device=dev)), 0).values
    occ = torch.empty(nnz, dtype=torch.int64, device=dev)
    occ[o1] = ar(nnz) - start
    assert nnz == 0 or int(occ.max()) < 8
    del k1, k1s, first, start
    # union entries: group by (pair, column, occurrence); at most one entry of each half per group
    k2 = ((gp_row[rows] * 65536 + lc) * 8 + occ) * 2 + half_row[rows]
    o2 = torch.argsort(k2)
    k2s = k2[o2]
    newg = torch.ones(nnz, dtype=torch.bool, device=dev)
    newg[1:] = (k2s[1:] >> 1) != (k2s[:-1] >> 1)
    uid = torch.cumsum(newg.to(torch.int64), 0) - 1                  # union id of every sorted entry
    nu = int(uid[-1]) + 1 if nnz else 0
    h_s = k2s & 1
    src = torch.full((nu, 2), -1, dtype=torch.int64, device=dev)      # CSR positions of the (A, B) values
    src[uid, h_s] = o2
    u_gp = torch.empty(nu, dtype=torch.int64, device=dev)
    u_lc = torch.empty(nu, dtype=torch.int64, device=dev)
    u_gp[uid] = k2s >> 20
    u_lc[uid] = (k2s >> 4) & 0xFFFF
    del k2, k2s, newg, uid, h_s, o1, o2, occ
    npairs = ntiles * (R // 2)
    ucount = torch.bincount(u_gp, minlength=npairs)
    # pairs of a tile by union length, descending -> pair index pi inside the tile
    pk = (ar(npairs) // (R // 2)) * (1 << 20) + ((1 << 20) - 1 - ucount)
    porder = torch.argsort(pk, stable=True)
    pi_of = torch.empty(npairs, dtype=torch.int64, device=dev)
    pi_of[porder] = ar(npairs) % (R // 2)
    tile_of_pair = ar(npairs) // (R // 2)
    blk_of_pair = tile_of_pair * 16 + (pi_of >> 2)
    steps = torch.zeros(nblk, dtype=torch.int64, device=dev)
    steps.scatter_reduce_(0, blk_of_pair, (ucount + 7) >> 3, reduce="amax")
    qptr = torch.zeros(nblk + 1, dtype=torch.int64, device=dev)
    torch.cumsum(steps * 32, 0, out=qptr[1:])
    nnzq = int(qptr[-1])
    if nnzq >= 2 ** 30 - 64:
        return None
    # row table
    qrow = torch.zeros(NT, dtype=torch.uint8, device=dev)
    qrow[(ar(NT) // R) * R + 2 * pi_of[gp_row] + half_row] = (ar(NT) % R).to(torch.uint8)
    # lane assignment: rank of every union entry inside its (pair, parity) class
    par = u_lc & 1
    k3 = u_gp * 2 + par
    o3 = torch.argsort(k3, stable=True)
    cnt3 = torch.bincount(k3, minlength=2 * npairs)
    st3 = torch.cumsum(cnt3, 0) - cnt3
    rk = torch.empty(nu, dtype=torch.int64, device=dev)
    rk[o3] = ar(nu) - st3[k3[o3]]
    cap = 4 * steps[blk_of_pair[u_gp]]                               # positions per lane class of the slot
    spill = rk >= cap
    cls = torch.where(spill, 1 - par, par)
    r = torch.where(spill, cnt3[u_gp * 2 + (1 - par)] + (rk - cap), rk)
    pos = qptr[blk_of_pair[u_gp]] + (r >> 2) * 32 + (pi_of[u_gp] & 3) * 8 + cls * 4 + (r & 3)
    # padding index per position: an own row of the slot's pair with the parity of the lane class
    p = ar(nnzq)
    blk = torch.repeat_interleave(ar(nblk), steps * 32)
    lane = (p - qptr[blk]) & 31
    tile_p = blk >> 4
    qa = tile_p * R + ((blk & 15) * 4 + (lane >> 3)) * 2              # table index of the slot's row A
    ra = qrow[qa].to(torch.int64)
    want = (lane >> 2) & 1
    cand = (ra & ~1) | want
    nrows_t = torch.clamp(n - tile_p * R, max=R)
    cand = torch.where(cand < nrows_t, cand, torch.where(ra < nrows_t, ra, torch.zeros_like(ra)))
    qcol = torch.zeros(nnzq + 64, dtype=torch.int16, device=dev)
    qcol[:nnzq] = cand.to(torch.int32).to(torch.int16)
    del p, blk, lane, tile_p, qa, ra, want, cand
    qcol[pos] = u_lc.to(torch.int32).to(torch.int16)
    qsrc = torch.full((nnzq + 64, 2), -1, dtype=torch.int32, device=dev)
    mine = torch.where(cls == 0, src[:, 0], src[:, 1])
    other = torch.where(cls == 0, src[:, 1], src[:, 0])
    qsrc[pos, 0] = mine.to(torch.int32)
    qsrc[pos, 1] = other.to(torch.int32)
    tile_tot = qptr[16::16] - qptr[:-1:16]
    qpad = torch.full((512 * ((ntiles + 31) // 32) + 4,), nnzq, dtype=torch.int32, device=dev)
    qpad[:nblk + 1] = qptr.to(torch.int32)
    return dict(qptr=qpad, qcol=qcol, qsrc=qsrc.reshape(-1).contiguous(), qrow=qrow, nnzq=nnzq, qnzmax=int(tile_tot.max()),
                q_unions=nu)


def wi_halo_lists(halo_ptr: torch.Tensor, halo_col: torch.Tensor, ntiles: int):
    """Halo id lists for the v5 kernel: every tile's list padded to a multiple of 4 ids (16-byte bulk copies) by repeating
    its last id (a valid row; 0 for empty lists is never read), offsets array padded to whole 32-tile chunks + 4."""
    dev = halo_ptr.device
    hp = halo_ptr.to(torch.int64)
    hlen = hp[1:] - hp[:-1]
    hlen4 = (hlen + 3) & ~3
    hptr4 = torch.zeros(ntiles + 1, dtype=torch.int64, device=dev)
    torch.cumsum(hlen4, 0, out=hptr4[1:])
    total4 = int(hptr4[-1])
    tile = torch.repeat_interleave(torch.arange(ntiles, device=dev, dtype=torch.int64), hlen4)
    k = torch.arange(total4, device=dev, dtype=torch.int64) - hptr4[tile]
    src = hp[tile] + torch.minimum(k, (hlen[tile] - 1).clamp_min(0))
    hcol4 = torch.zeros(total4 + 4, dtype=torch.int32, device=dev)
    if total4:
        hcol4[:total4] = halo_col[src].to(torch.int32)
    hpad = torch.full((32 * ((ntiles + 31) // 32) + 4,), total4, dtype=torch.int32, device=dev)
    hpad[:ntiles + 1] = hptr4.to(torch.int32)
    return dict(hptr=hpad, hcol=hcol4, hmax=int(hlen4.max()) if ntiles else 0)


class GraphStructure:
    """rowptr[n+1], col[nnz], eid[nnz] (int32) with nnz = 2M; column indices ascending inside a row.

    If ``perm`` (new position -> old index) is given, rows AND columns live in the permuted numbering: vectors must
    be permuted (``x[perm]``) before and un-permuted (``y[inv]``) after a product; the solvers do this once per solve."""

    def __init__(self, idx: torch.Tensor, n: int, perm: torch.Tensor = None):
        if not idx.is_cuda:
            raise RuntimeError("GraphStructure: edge index must be a CUDA tensor (no CPU fallback exists)")
        if idx.dim() != 2 or idx.shape[0] != 2:
            raise ValueError("edge index must be [2, M]")
        idx = idx.to(torch.int64)
        self.perm = self.inv = None
        self.morton_pos = None
        if perm is not None:
            perm = perm.to(device=idx.device, dtype=torch.int64)
            # balance the tiled kernel's sub-warps: inside every tile of TILE_ROWS consecutive rows, order the rows by
            # their number of nonzeros (descending), so rows processed together have (nearly) equal length
            deg = torch.bincount(idx.reshape(-1), minlength=int(n))
            tile_id = torch.arange(perm.numel(), device=idx.device, dtype=torch.int64) // self.TILE_ROWS
            key = (tile_id << 20) | ((1 << 20) - 1 - deg[perm].clamp_max((1 << 20) - 1))
            order = torch.argsort(key, stable=True)
            perm = perm[order]
            self.morton_pos = order                     # position of each final row in the spatial order
            self.perm = perm.contiguous()
            self.inv = torch.empty_like(self.perm)
            self.inv[self.perm] = torch.arange(self.perm.numel(), device=idx.device)
            idx = self.inv[idx]
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
        self.perm32 = None if self.perm is None else self.perm.to(torch.int32).contiguous()
        self.tiles = None
        self._tiles_tried = False
        self.build_tiles()   # now: it may reorder the entries inside rows, which must happen before any value build

    @classmethod
    def from_rows(cls, rowptr: torch.Tensor, col: torch.Tensor, n_rows: int, n_cols: int):
        """Rank-local structure of a ROW PARTITION (distributed.PartitionedGraph): ``n_rows`` rows this rank owns, entries of
        BOTH directions of every incident edge already present (``rowptr`` [n_rows + 1], ``col`` [nnz] in the rank's extended
        numbering: own rows 0 .. n_rows - 1 first, halo rows n_rows .. n_cols - 1 after).  Vectors the SpMM reads have ``n_cols``
        rows, vectors it writes ``n_rows``.  ``eid`` is the identity: per-entry values are passed where the square structure
        takes per-edge values (``d2csr`` gathers through ``eid``, so the in-row reordering of ``build_tiles`` is followed)."""
        if not col.is_cuda:
            raise RuntimeError("GraphStructure.from_rows: arrays must be CUDA tensors (no CPU fallback exists)")
        st = object.__new__(cls)
        st.perm = st.inv = st.perm32 = st.morton_pos = None
        st.n, st.n_cols = int(n_rows), int(n_cols)
        st.nnz = int(col.numel())
        st.m = st.nnz
        st.device = col.device
        st.rowptr = rowptr.to(torch.int32).contiguous()
        st.col = col.to(torch.int32).contiguous()
        st.eid = torch.arange(st.nnz, dtype=torch.int32, device=st.device)
        st._d2 = {}
        st._dot_ws = None
        st._upper_pos = None
        st.tiles = None
        st._tiles_tried = False
        st.build_tiles()
        return st

    # -- tile-compacted structure for the v2 SpMM kernel (csrc/lap_spmm_tiled.cu) ----------------------------------------
    TILE_ROWS = 128
    TILED_SMEM_LIMIT = 200 * 1024

    def build_tiles(self):
        """Per tile of TILE_ROWS consecutive rows: the sorted list of distinct out-of-tile columns (halo) and a 16-bit
        tile-local column index per nonzero.  Setup only (torch sort/unique), once per graph."""
        if self._tiles_tried:
            return self.tiles
        self._tiles_tried = True
        R, n, dev = self.TILE_ROWS, self.n, self.device
        ntiles = (n + R - 1) // R
        rowlen = (self.rowptr[1:] - self.rowptr[:-1]).to(torch.int64)
        rows = torch.repeat_interleave(torch.arange(n, device=dev, dtype=torch.int64), rowlen)
        tile = rows // R
        col = self.col.to(torch.int64)
        lo = tile * R
        own = (col >= lo) & (col < lo + R)
        lcol = torch.where(own, col - lo, torch.zeros_like(col))
        out = ~own
        key = (tile[out] << 32) | col[out]
        ukey, inverse = torch.unique(key, sorted=True, return_inverse=True)
        bounds = torch.arange(ntiles + 1, device=dev, dtype=torch.int64) << 32
        halo_ptr = torch.searchsorted(ukey, bounds)
        lcol[out] = R + inverse - halo_ptr[tile[out]]
        # bank-conflict control for 64-byte X rows in shared memory: two nonzeros handled by the same quarter-warp hit
        # different bank halves iff their local indices have different parity.  Rows in even sub-warp slots list their
        # even-parity entries first, rows in odd slots their odd-parity entries first (entry order inside a row is free).
        flip = ((lcol & 1) ^ (rows & 1)).to(torch.int64)
        order = torch.argsort((rows << 1) | flip, stable=True)
        if not bool((order[1:] > order[:-1]).all()):
            self.col = self.col[order].contiguous()
            self.eid = self.eid[order].contiguous()
            lcol = lcol[order]
            self._d2 = {}
            self._upper_pos = None
        hlen = halo_ptr[1:] - halo_ptr[:-1]
        tstart = self.rowptr[torch.arange(0, n, R, device=dev)].to(torch.int64)
        tend = torch.cat([tstart[1:], self.rowptr[-1:].to(torch.int64)])
        lmax = int(R + (hlen.max() if hlen.numel() else 0))
        nzmax = int((tend - tstart).max())
        if lmax > 65535:
            return None
        lcol16 = torch.zeros(self.nnz + 8, dtype=torch.int16, device=dev)
        lcol16[:self.nnz] = lcol.to(torch.int32).to(torch.int16)          # two's-complement wrap == uint16 bit pattern
        self.tiles = dict(lcol=lcol16, halo_ptr=halo_ptr.to(torch.int32).contiguous(),
                          halo_col=torch.cat([(ukey & 0xFFFFFFFF).to(torch.int32),
                                              torch.zeros(1, dtype=torch.int32, device=dev)]).contiguous(),
                          lmax=lmax, nzmax=nzmax, rows=R, halo_total=int(ukey.numel()))
        padded = padded_streams(self.rowptr, lcol16[:self.nnz], n, R)
        if padded is not None:
            self.tiles.update(padded)
        wi = wi_streams(self.rowptr, lcol16[:self.nnz], n, R)
        if wi is not None:
            self.tiles.update(wi)
            self.tiles.update(wi_halo_lists(self.tiles["halo_ptr"], self.tiles["halo_col"], ntiles))
        return self.tiles

    def _value_layout(self, a: torch.Tensor, kind: str) -> torch.Tensor:
        """``a`` (CSR entry order, one bandwidth) in the stream layout of a pipelined kernel (``kind``: "pad" = v4 padded
        rows, "wi" = v5 warp-interleaved).  Cached per value array: the cache holds a reference to ``a`` itself (so its
        address cannot be recycled while the entry lives) and matches on (address, version), which ``a.detach()`` shares."""
        cache = self.__dict__.setdefault("_layout_cache", [])
        for k, ref, ver, out in cache:
            if k == kind and ref.data_ptr() == a.data_ptr() and ver == a._version and ref.dtype == a.dtype and ref.numel() == a.numel():
                log = self.__dict__.get("_layout_log")
                if log is not None:
                    log.append(out)
                return out
        t = self.tiles
        sfx = _lib.suffix(a.dtype)
        if kind == "pad":
            out = torch.empty(t["nnzp"] + 8, dtype=a.dtype, device=a.device)
            out[t["nnzp"]:].zero_()
            _lib.call("mgp_lap_pad_values_" + sfx, ptr(self.rowptr), ptr(t["prowptr"]), ptr(a), c_int64(self.n), ptr(out), stream())
        elif kind == "pair":
            out = torch.empty(t["qsrc"].numel(), dtype=a.dtype, device=a.device)
            _lib.call("mgp_lap_pair_values_" + sfx, ptr(t["qsrc"]), ptr(a), c_int64(out.numel()), ptr(out), stream())
        else:
            out = torch.empty(t["nnzw"] + 64, dtype=a.dtype, device=a.device)
            out[t["nnzw"]:].zero_()
            _lib.call("mgp_lap_wi_values_" + sfx, ptr(self.rowptr), ptr(t["wptr"]), ptr(a), c_int64(self.n), ptr(out), stream())
        cache.append((kind, a.detach(), a._version, out))
        del cache[:-4]
        log = self.__dict__.get("_layout_log")
        if log is not None:
            log.append(out)
        return out

    def padded_values(self, a: torch.Tensor) -> torch.Tensor:
        return self._value_layout(a, "pad")

    def wi_values(self, a: torch.Tensor) -> torch.Tensor:
        if self.__dict__.get("_aw_persistent") is not None:
            return self._aw_persistent     # the owner keeps ONE buffer alive across bandwidths (captured CUDA graphs point at it)
        return self._value_layout(a, "wi")

    def pair_tiles(self):
        """The paired-row streams (``pair_streams``) of this structure, built on first use (fp32 SpMM with whole 64-byte
        rows of right-hand sides); None when the tile structure does not exist."""
        t = self.build_tiles()
        if t is None or "wptr" not in t:
            return None
        if "qptr" not in t and not self.__dict__.get("_pair_tried"):
            self._pair_tried = True
            if (self.TILE_ROWS + t["hmax"]) * 64 + t["wnzmax"] * 6 > 200 * 1024:
                return None                  # a tile this wide does not fit the kernel's ring in either layout: do not build the streams
            lc = t["lcol"][:self.nnz]
            q = pair_streams(self.rowptr, lc, self.n, pair_matching(self.rowptr, lc, self.n, self.TILE_ROWS), self.TILE_ROWS)
            if q is not None:
                t.update(q)
        return t if "qptr" in t else None

    def pair_values(self, a: torch.Tensor) -> torch.Tensor:
        if self.__dict__.get("_aq_persistent") is not None:
            return self._aq_persistent     # see wi_values
        return self._value_layout(a, "pair")

    def tiled_ok(self, dtype, cw: int) -> bool:
        """Does the widest pass for ``cw`` columns of ``dtype`` fit the shared-memory budget of the tiled kernel?"""
        t = self.build_tiles()
        if t is None:
            return False
        w = 4 if dtype == torch.float32 else 8
        cwp = 1
        while cwp < min(cw, 16):
            cwp *= 2
        if cw > 16:
            cwp = 32 if (cw % (4 if w == 4 else 2)) else 16
        lmax = (t["lmax"] + 3) & ~3
        nzcap = (t["nzmax"] + 15) & ~7
        return lmax * cwp * w + nzcap * (w + 2) + (t["rows"] + 4) * 4 + 2 * t["rows"] * w + 16 <= self.TILED_SMEM_LIMIT

    # -- cached helpers -------------------------------------------------------------------------------------------
    def d2csr(self, val: torch.Tensor) -> torch.Tensor:
        """Per-directed-entry copy of the per-edge squared distances (cached per (storage, dtype))."""
        key = (val.data_ptr(), val.dtype, val._version)
        hit = self._d2.get(key)
        if hit is not None:
            hit = hit[0]
        if hit is None:
            v = val.detach().reshape(-1).contiguous()
            out = torch.empty(self.nnz, dtype=v.dtype, device=self.device)
            _lib.call("mgp_gather_edge_" + _lib.suffix(v.dtype), ptr(v), ptr(self.eid), c_int64(self.nnz), ptr(out), stream())
            # keep only the latest; the entry holds ``val`` itself so its address cannot be recycled (by the caching
            # allocator, for another edge-value tensor of the same size) while the key is live
            self._d2 = {key: (out, val)}
            hit = out
        return hit

    def dot_ws(self) -> torch.Tensor:
        if self._dot_ws is None:
            nb = _lib.query("mgp_lap_spmm_dot_ws_bytes", c_int64(self.n), c_int32(32))
            self._dot_ws = torch.zeros(nb, dtype=torch.uint8, device=self.device)
        return self._dot_ws

    def upper_pos(self) -> torch.Tensor:
        """Position in the CSR arrays of one directed copy of every undirected edge e -- maps per-entry arrays back to
        the reference's per-edge arrays (laplacian_triu etc.; the entries are symmetric so either copy serves)."""
        if self._upper_pos is None:
            out = torch.empty(self.m, dtype=torch.int64, device=self.device)
            out[self.eid.to(torch.int64)] = torch.arange(self.nnz, device=self.device)
            self._upper_pos = out
        return self._upper_pos

    # vectors between the caller's order and the structure's (permuted) order
    def to_internal(self, v: torch.Tensor) -> torch.Tensor:
        return v if self.perm is None else v.index_select(0, self.perm)

    def to_external(self, v: torch.Tensor) -> torch.Tensor:
        return v if self.inv is None else v.index_select(0, self.inv)


_STRUCT_ATTR = "_mgp_b200_structure"


def structure_for(idx: torch.Tensor, n: int) -> GraphStructure:
    """Build (once) and cache the GraphStructure on the edge-index tensor object."""
    st = getattr(idx, _STRUCT_ATTR, None)
    if st is None or st.n != int(n) or st.m != int(idx.shape[1]) or st.device != idx.device:
        st = GraphStructure(idx, n, perm=getattr(idx, _PERM_ATTR, None))
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
    a = torch.zeros(st.nnz + 8, dtype=dt, device=dev)[:st.nnz]   # 8 entries of slack: the tiled kernel streams 128-bit chunks
    _lib.call("mgp_lap_values_" + _lib.suffix(dt), ptr(st.rowptr), ptr(st.col), ptr(d2csr), c_int64(st.n), ptr(eps_t),
              c_int32(1 if self_loops else 0), ptr(deg_un), ptr(deg), ptr(diag), ptr(a), stream())
    return deg_un, deg, diag, a


# ---- SpMM ---------------------------------------------------------------------------------------------------------
LAST_SPMM_KERNEL = None  # name of the kernel the most recent lap_spmm call launched (bench.py reports it)
SPMM_KERNEL = "auto"   # "auto" | "csr" | "tiled" | "pipe" | "wi" | "wp" | "spmv"  (tests force each; "auto": wp (fp32, whole 64-byte rows), else wi, else pipe, else tiled, else csr; one column: spmv)


# "auto" may take the paired-row walk of the warp-interleaved kernel (MGP_PAIR_WALK=0 turns it off)
PAIR_WALK = os.environ.get("MGP_PAIR_WALK", "1") != "0"


def _note_kernel(name):
    global LAST_SPMM_KERNEL
    LAST_SPMM_KERNEL = name


def lap_spmm(st: GraphStructure, a, diag, x, shift=None, pre=None, post=None, out=None, dot_with=None, dot_out=None,
             x_external=False, y_external=False, peer_x=None, peer_sync=None, peer_ext=None, done_flag=None, ep_coef=None,
             ep_add=None):
    """``_lap_spmm`` plus the wrappers' epilogue algebra: Y <- ep_add + ep_coef * Y (``ep_coef``: device scalar tensor, ``ep_add``:
    [n, C] like X; a requested dot product is taken with the scaled Y, and excludes ``ep_add``).  The warp-interleaved kernel
    applies it inside the launch (mgp_wi_ext.ep_coef / ep_add); for every other kernel it is two elementwise passes here."""
    if ep_coef is None and ep_add is None:
        return _lap_spmm(st, a, diag, x, shift, pre, post, out, dot_with, dot_out, x_external, y_external, peer_x, peer_sync,
                         peer_ext, done_flag)
    if ep_add is not None and (dot_out is not None or ep_coef is None):
        raise ValueError("lap_spmm: ep_add needs ep_coef and excludes the dot epilogue")
    handled = []
    y = _lap_spmm(st, a, diag, x, shift, pre, post, out, dot_with, dot_out, x_external, y_external, peer_x, peer_sync, peer_ext,
                  done_flag, _ep=(ep_coef, ep_add, handled))
    if not handled:
        cf = ep_coef.to(y.dtype)
        y.mul_(cf)
        if ep_add is not None:
            y.add_(ep_add[:, :y.shape[1]])
        if dot_out is not None:
            dot_out.mul_(cf)
    return y


def _lap_spmm(st: GraphStructure, a, diag, x, shift=None, pre=None, post=None, out=None, dot_with=None, dot_out=None,
              x_external=False, y_external=False, peer_x=None, peer_sync=None, peer_ext=None, done_flag=None, _ep=None):
    """Y = post .* ((diag + shift) .* (pre .* X) - A (pre .* X)).  ``x``: [n, C] CUDA tensor with unit column stride.

    ``peer_x``: int64 device tensor of every rank's X base pointer (row-partitioned multi-GPU, see distributed.PeerCG);
    ``peer_sync`` = (rank, flag pointer table, epoch scalar) fuses the cross-GPU barrier into the launch;
    ``peer_ext`` = (rank, epoch scalar, ``_lib.wi_ext(...)``) selects the extended hooks of mgp_lap_spmm_wi_ex instead (lazy
    wait / publish-at-end flags, shipping of the dot partials); ``done_flag``: device scalar that turns the launch into a
    no-op when non-zero (warp-interleaved kernel only; other kernels ignore it and compute).
    ``a, diag, pre, post`` are in the structure's row order.  ``x_external`` / ``y_external``: X (and ``dot_with``) / Y are
    in the caller's row order (only matters when the structure is internally permuted)."""
    if x.dim() != 2 or x.shape[0] < st.n:   # a row-partitioned structure reads [own rows | halo rows]: more rows than it writes
        raise ValueError(f"lap_spmm: expected rhs of shape [{st.n}, C], got {tuple(x.shape)}")
    if x.stride(1) != 1:
        x = x.contiguous()
    dt = x.dtype
    if a.dtype != dt or diag.dtype != dt:
        raise TypeError(f"lap_spmm: value dtype {a.dtype} does not match rhs dtype {dt}")
    c = int(x.shape[1])
    if st.perm is None:
        x_external = y_external = False
    shift_t = None if shift is None else _device_scalar(shift, dt, x.device)
    ws = st.dot_ws() if dot_out is not None else None
    sfx = _lib.suffix(dt)
    slack_ok = a.untyped_storage().nbytes() >= (a.storage_offset() + st.nnz + 8) * a.element_size()
    # measured on B200 (profiles/): the tile-compacted kernels win from 4 columns up; for 1-3 columns the CSR sub-warp
    # kernel (X served from L1/L2) is faster
    use_tiled = SPMM_KERNEL not in ("csr", "spmv") and slack_ok and st.tiled_ok(dt, c) and (c >= 4 or SPMM_KERNEL in ("tiled", "pipe", "wi", "wp"))
    if SPMM_KERNEL in ("tiled", "pipe", "wi", "wp") and not use_tiled:
        raise RuntimeError("lap_spmm: tiled kernel requested but the tile structure does not fit in shared memory")
    if use_tiled:
        t = st.tiles
        if out is None:
            out = torch.empty((st.n, c), dtype=dt, device=x.device)
        if peer_x is not None and not (pre is None and "wptr" in t):
            raise RuntimeError("lap_spmm: peer-memory halo reads need the warp-interleaved kernel (no pre scaling)")
        if pre is None and "wptr" in t and (SPMM_KERNEL in ("auto", "wi", "wp") or peer_x is not None):
            # paired-row walk (v6): fp32, whole 64-byte rows of right-hand sides, single GPU
            # (with peer memory: only when the caller's hook block names the row table, i.e. keeps the paired value stream alive)
            ext_pair = peer_ext is not None and bool(peer_ext[2].pair_rows)
            pair = (SPMM_KERNEL in ("auto", "wp") and PAIR_WALK and dt == torch.float32 and c % 16 == 0
                    and ((peer_x is None and peer_ext is None) or ext_pair)
                    and hasattr(st, "pair_tiles") and st.pair_tiles() is not None)
            if ext_pair and not pair:
                raise RuntimeError("lap_spmm: the hook block asks for the paired-row walk but this call does not qualify")
            if SPMM_KERNEL == "wp" and not pair:
                raise RuntimeError("lap_spmm: paired-row kernel requested but this call does not qualify (fp32, multiples of 16 "
                                   "columns, no peer memory)")
            aw = st.pair_values(a) if pair else st.wi_values(a)
            sptr, scol, snzmax = (t["qptr"], t["qcol"], t["qnzmax"]) if pair else (t["wptr"], t["wcol"], t["wnzmax"])
            hcol = t["hcol_peer"] if peer_x is not None else t["hcol"]
            ep_here = _ep is not None and peer_ext is None and not (x_external or y_external)
            if peer_ext is not None or done_flag is not None or ep_here or pair:
                import ctypes
                if peer_ext is not None:
                    ext = peer_ext[2]
                elif ep_here:
                    cf = _ep[0] if _ep[0].dtype == dt else _ep[0].to(dt)
                    ext = _lib.wi_ext(done_flag=done_flag, ep_coef=cf, ep_add=_ep[1] is not None, pair_rows=t["qrow"] if pair else None)
                    if _ep[1] is not None:
                        dot_with = _ep[1]
                    _ep[2].append(True)
                else:
                    ext = _lib.wi_ext(done_flag=done_flag, pair_rows=t["qrow"] if pair else None)
                rc = _lib.call_rc("mgp_lap_spmm_wi_ex_" + sfx, ptr(sptr), ptr(scol), ptr(aw), ptr(diag),
                                  ptr(t["hptr"]), ptr(hcol), c_int32(t["rows"]), c_int32(t["rows"] + t["hmax"]),
                                  c_int32(snzmax), c_int32(t["hmax"]), ptr(shift_t), ptr(post),
                                  ptr(st.perm32 if x_external else None),
                                  ptr(st.perm32 if y_external else None), ptr(x), c_int64(x.stride(0)), ptr(out),
                                  c_int64(out.stride(0)), c_int64(st.n), c_int32(c), ptr(dot_with), ptr(dot_out), ptr(ws),
                                  ptr(peer_x), c_int32(0 if peer_x is None else int(peer_x.numel())),
                                  c_int32(0 if peer_ext is None else int(peer_ext[0])),
                                  ptr(None if peer_ext is None else peer_ext[1]), ctypes.byref(ext), stream())
            else:
                rc = _lib.call_rc("mgp_lap_spmm_wi_" + sfx, ptr(t["wptr"]), ptr(t["wcol"]), ptr(aw), ptr(diag),
                                  ptr(t["hptr"]), ptr(hcol), c_int32(t["rows"]), c_int32(t["rows"] + t["hmax"]),
                                  c_int32(t["wnzmax"]), c_int32(t["hmax"]), ptr(shift_t), ptr(post),
                                  ptr(st.perm32 if x_external else None),
                                  ptr(st.perm32 if y_external else None), ptr(x), c_int64(x.stride(0)), ptr(out),
                                  c_int64(out.stride(0)), c_int64(st.n), c_int32(c), ptr(dot_with), ptr(dot_out), ptr(ws),
                                  ptr(peer_x), c_int32(0 if peer_x is None else int(peer_x.numel())),
                                  c_int32(0 if peer_sync is None else int(peer_sync[0])), ptr(None if peer_sync is None else peer_sync[1]),
                                  ptr(None if peer_sync is None else peer_sync[2]), stream())
            if rc == 0:
                _note_kernel("lap_spmm_wi_kernel<pair>" if pair else "lap_spmm_wi_kernel")
                return out
            if _ep is not None and _ep[2]:          # not launched: the epilogue algebra falls to the caller's elementwise passes
                _ep[2].clear()
                if _ep[1] is not None:
                    dot_with = None
            if rc != _lib.MGP_EUNSUPPORTED or peer_x is not None:
                raise RuntimeError(f"mgp_lap_spmm_wi_{sfx} failed ({rc}): {_lib.last_error()}")
        if SPMM_KERNEL in ("wi", "wp"):
            raise RuntimeError("lap_spmm: warp-interleaved kernel requested but this call does not qualify (pre scaling, "
                               "column count not a multiple of one 64-byte row, alignment or shared memory)")
        if pre is None and "prowptr" in t and SPMM_KERNEL in ("auto", "pipe"):
            ap = st.padded_values(a)
            rc = _lib.call_rc("mgp_lap_spmm_pipe_" + sfx, ptr(t["prowptr"]), ptr(t["plcol"]), ptr(ap), ptr(diag),
                              ptr(t["halo_ptr"]), ptr(t["halo_col"]), c_int32(t["rows"]), c_int32(t["lmax"]),
                              c_int32(t["pnzmax"]), ptr(shift_t), ptr(post), ptr(st.perm32 if x_external else None),
                              ptr(st.perm32 if y_external else None), ptr(x), c_int64(x.stride(0)), ptr(out),
                              c_int64(out.stride(0)), c_int64(st.n), c_int32(c), ptr(dot_with), ptr(dot_out), ptr(ws), stream())
            if rc == 0:
                _note_kernel("lap_spmm_pipe_kernel")
                return out
            if rc != _lib.MGP_EUNSUPPORTED:
                raise RuntimeError(f"mgp_lap_spmm_pipe_{sfx} failed ({rc}): {_lib.last_error()}")
        if SPMM_KERNEL == "pipe":
            raise RuntimeError("lap_spmm: pipelined kernel requested but this call does not qualify (pre scaling, column "
                               "count / alignment or shared memory)")
        rc = _lib.call_rc("mgp_lap_spmm_tiled_" + sfx, ptr(st.rowptr), ptr(t["lcol"]), ptr(a), ptr(diag), ptr(t["halo_ptr"]),
                          ptr(t["halo_col"]), c_int32(t["rows"]), c_int32(t["lmax"]), c_int32(t["nzmax"]), ptr(shift_t),
                          ptr(pre), ptr(post), ptr(st.perm32 if x_external else None),
                          ptr(st.perm32 if y_external else None), ptr(x), c_int64(x.stride(0)), ptr(out),
                          c_int64(out.stride(0)), c_int64(st.n), c_int32(c), ptr(dot_with), ptr(dot_out), ptr(ws), stream())
        if rc == 0:
            _note_kernel("lap_spmm_tiled_kernel")
            return out
        if rc != _lib.MGP_EUNSUPPORTED or SPMM_KERNEL == "tiled":
            raise RuntimeError(f"mgp_lap_spmm_tiled_{sfx} failed ({rc}): {_lib.last_error()}")
        # a pass wider than estimated (unaligned leading dimension): the CSR kernel handles it
    if peer_x is not None:
        raise RuntimeError("lap_spmm: peer-memory halo reads requested but the tile structure does not fit this call")
    # one column, no pre scaling: the streamed single-column kernel on the tile streams (lap_spmv_tile.cu)
    if c == 1 and pre is None and SPMM_KERNEL in ("auto", "spmv") and slack_ok:
        t = st.build_tiles()
        if t is not None and "wptr" in t:
            if out is None:
                out = torch.empty((st.n, 1), dtype=dt, device=x.device)
            aw = st.wi_values(a)
            rc = _lib.call_rc("mgp_lap_spmv_tile_" + sfx, ptr(t["wptr"]), ptr(t["wcol"]), ptr(aw), ptr(diag), ptr(t["hptr"]),
                              ptr(t["hcol"]), c_int32(t["rows"]), c_int32(t["hmax"]), ptr(shift_t), ptr(post),
                              ptr(st.perm32 if x_external else None), ptr(st.perm32 if y_external else None), ptr(x),
                              c_int64(x.stride(0)), ptr(out), c_int64(out.stride(0)), c_int64(st.n), ptr(dot_with), ptr(dot_out),
                              ptr(ws), stream())
            if rc == 0:
                _note_kernel("lap_spmv_tile_kernel")
                return out
            if rc != _lib.MGP_EUNSUPPORTED:
                raise RuntimeError(f"mgp_lap_spmv_tile_{sfx} failed ({rc}): {_lib.last_error()}")
    if SPMM_KERNEL == "spmv":
        raise RuntimeError("lap_spmm: single-column tile kernel requested but this call does not qualify")
    # v1 CSR kernel works in the structure's order only
    if x_external:
        x = x.index_select(0, st.perm)
        if dot_with is not None:
            dot_with = dot_with.index_select(0, st.perm)
    y = torch.empty((st.n, c), dtype=dt, device=x.device) if (out is None or y_external) else out
    _note_kernel("lap_spmm_csr_kernel")
    _lib.call("mgp_lap_spmm_" + sfx, ptr(st.rowptr), ptr(st.col), ptr(a), ptr(diag), ptr(shift_t),
              ptr(pre), ptr(post), ptr(x), c_int64(x.stride(0)), ptr(y), c_int64(y.stride(0)), c_int64(st.n),
              c_int32(c), ptr(dot_with), ptr(dot_out), ptr(ws), stream())
    if y_external:
        if out is None:
            return y.index_select(0, st.inv)
        out.copy_(y.index_select(0, st.inv))
        return out
    return y


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
