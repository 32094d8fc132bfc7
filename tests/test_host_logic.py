"""Host-side index plumbing that needs no GPU: the padded entry streams of the pipelined SpMM kernel."""
import torch


def test_padded_streams_layout():
    from manifold_gp_b200.graph import padded_streams
    g = torch.Generator().manual_seed(0)
    for n, R, maxlen in ((300, 128, 9), (128, 128, 40), (1000, 128, 3), (5, 128, 7)):
        rowlen = torch.randint(0, maxlen, (n,), generator=g)
        rowptr = torch.zeros(n + 1, dtype=torch.int32)
        rowptr[1:] = torch.cumsum(rowlen, 0)
        nnz = int(rowptr[-1])
        lcol = torch.randint(0, 400, (nnz,), generator=g).to(torch.int16)
        t = padded_streams(rowptr, lcol, n, R)
        pr, pl = t["prowptr"].long(), t["plcol"]
        assert pl.numel() == t["nnzp"] + 8 and t["nnzp"] % 8 == 0 and t["pnzmax"] % 8 == 0
        assert all(int(pr[i]) % 8 == 0 for i in range(0, n, R))             # tiles start at multiples of 8 entries
        assert bool(((pr[1:] - pr[:-1]) % 4 == 0).all())                     # rows hold multiples of 4 entries
        for r in range(n):
            L, q0, q1 = int(rowlen[r]), int(pr[r]), int(pr[r + 1])
            assert torch.equal(pl[q0:q0 + L], lcol[int(rowptr[r]):int(rowptr[r]) + L])
            assert bool((pl[q0 + L:q1] == r % R).all())                      # padding points at the row itself


def test_wi_streams_layout():
    """Warp-interleaved streams: position wptr[b] + 32 t + lane <-> nonzero 4t + (lane & 3) of row 128 tile + 8 w + lane // 4."""
    from manifold_gp_b200.graph import wi_streams
    g = torch.Generator().manual_seed(1)
    R = 128
    for n, maxlen in ((300, 9), (128, 40), (1000, 3), (5, 7), (257, 50)):
        rowlen = torch.randint(0, maxlen, (n,), generator=g)
        rowptr = torch.zeros(n + 1, dtype=torch.int32)
        rowptr[1:] = torch.cumsum(rowlen, 0)
        nnz = int(rowptr[-1])
        lcol = torch.randint(0, 400, (nnz,), generator=g).to(torch.int16)
        t = wi_streams(rowptr, lcol, n, R)
        wptr, wcol = t["wptr"].long(), t["wcol"]
        ntiles = (n + R - 1) // R
        assert wptr.numel() == 512 * ((ntiles + 31) // 32) + 4          # padded to whole 32-tile metadata chunks
        assert bool((wptr[16 * ntiles:] == t["nnzw"]).all())
        assert wcol.numel() == t["nnzw"] + 64 and t["wnzmax"] % 32 == 0
        seen = 0
        for blk in range(ntiles * 16):
            o, size = int(wptr[blk]), int(wptr[blk + 1] - wptr[blk])
            assert size % 32 == 0
            steps = size // 32
            for lane in range(32):
                rl = (blk % 16) * 8 + lane // 4
                row = (blk // 16) * R + rl
                L = int(rowlen[row]) if row < n else 0
                assert (L + 3) // 4 <= steps
                for tt in range(steps):
                    e, pos = 4 * tt + (lane & 3), o + 32 * tt + lane
                    if e < L:
                        assert int(wcol[pos]) == int(lcol[int(rowptr[row]) + e])
                        seen += 1
                    else:   # padding: neighbour row (opposite parity) when it exists, 0 for rows beyond n
                        exp = 0 if row >= n else ((rl ^ 1) if (blk // 16) * R + (rl ^ 1) < n else rl)
                        assert int(wcol[pos]) == exp
        assert seen == nnz


def test_wi_halo_lists_padding():
    from manifold_gp_b200.graph import wi_halo_lists
    g = torch.Generator().manual_seed(2)
    for ntiles in (1, 3, 40):
        hlen = torch.randint(0, 11, (ntiles,), generator=g)
        hptr = torch.zeros(ntiles + 1, dtype=torch.int32)
        hptr[1:] = torch.cumsum(hlen, 0)
        hcol = torch.randint(0, 10000, (int(hptr[-1]) + 1,), generator=g).to(torch.int32)
        t = wi_halo_lists(hptr, hcol, ntiles)
        hp, hc = t["hptr"].long(), t["hcol"]
        assert hp.numel() == 32 * ((ntiles + 31) // 32) + 4 and t["hmax"] % 4 == 0
        assert bool((hp[:ntiles + 1] % 4 == 0).all()) and bool((hp[ntiles:] == hp[ntiles]).all())
        for k in range(ntiles):
            L, a, b = int(hlen[k]), int(hp[k]), int(hp[k + 1])
            assert b - a == (L + 3) // 4 * 4 and b - a <= t["hmax"]
            assert torch.equal(hc[a:a + L], hcol[int(hptr[k]):int(hptr[k]) + L])
            if L:
                assert bool((hc[a + L:b] == hcol[int(hptr[k]) + L - 1]).all())   # padding repeats the last (valid) id


def test_knn_graph_file_round_trip(tmp_path):
    """graph.save_knn_graph / load_knn_graph: the edge list, values and the attached row order survive,
    and a file built from other points or another k is refused."""
    import pytest
    from manifold_gp_b200 import graph
    g = torch.Generator().manual_seed(0)
    n, k = 500, 6
    x = torch.randn(n, 3, generator=g)
    idx = torch.stack([torch.randint(0, n - 1, (900,), generator=g), torch.randint(0, n, (900,), generator=g)])
    val = torch.rand(900, generator=g)
    perm = torch.randperm(n, generator=g)
    graph.attach_permutation(idx, perm)
    path = str(tmp_path / "g.pt")
    graph.save_knn_graph(path, idx, val, n, k, x=x)
    idx2, val2 = graph.load_knn_graph(path, "cpu", x=x, k=k)
    assert idx2.dtype == torch.int64 and torch.equal(idx2, idx) and torch.equal(val2, val)
    assert torch.equal(getattr(idx2, graph._PERM_ATTR), perm)
    with pytest.raises(ValueError):
        graph.load_knn_graph(path, "cpu", x=x + 1.0, k=k)
    with pytest.raises(ValueError):
        graph.load_knn_graph(path, "cpu", x=x, k=k + 1)
    with pytest.raises(ValueError):
        graph.load_knn_graph(path, "cpu", x=x[:-1], k=k)


_FAKE_LINEAR_OPERATOR = '''
import torch


class LinearOperator:
    """Stand-in for linear_operator.LinearOperator: representation / shape plumbing + EAGER solver entry points that must never run."""

    def __init__(self, *args, **kwargs):
        self._args, self._kwargs = args, kwargs

    def representation(self):
        out = []
        for a in list(self._args) + list(self._kwargs.values()):
            if torch.is_tensor(a):
                out.append(a)
            elif isinstance(a, LinearOperator):
                out += list(a.representation())
        return tuple(out)

    @property
    def shape(self):
        return torch.Size(self._size())

    @property
    def dtype(self):
        return next(t.dtype for t in self.representation() if t.is_floating_point())

    @property
    def device(self):
        return self.representation()[0].device

    def matmul(self, rhs):
        return self._matmul(rhs)

    def solve(self, *a, **k):
        raise AssertionError("linear_operator's eager solve was reached")

    def inv_quad_logdet(self, *a, **k):
        raise AssertionError("linear_operator's eager inv_quad_logdet was reached")

    def diagonalization(self, *a, **k):
        raise AssertionError("linear_operator's eager diagonalization was reached")
'''

_REAL_BASE_SCRIPT = '''
import torch, linear_operator
import manifold_gp_b200 as mgp
from manifold_gp_b200 import solvers
from manifold_gp_b200._compat import linear_operator as compat
assert compat.HAVE_LINEAR_OPERATOR and issubclass(compat.LinearOperator, linear_operator.LinearOperator)
calls = []
solvers.solve = lambda op, rhs: calls.append(("solve", type(op).__name__)) or rhs
solvers.inv_quad_logdet = lambda op, **kw: calls.append(("iql", type(op).__name__)) or (torch.zeros(()), torch.zeros(()))
solvers.diagonalization = lambda op, method=None: calls.append(("diag", type(op).__name__)) or (torch.zeros(3), compat.DenseEigenvectors(torch.zeros(6, 3)))
n = 6
idx = torch.tensor([[0, 1, 2, 3, 4], [1, 2, 3, 4, 5]])
val = torch.rand(5)
lap = mgp.GraphLaplacianOperator(val, idx, n, torch.tensor([[0.5]]), "symmetric", True)
prec = mgp.PrecisionMaternOperator(lap, 2, torch.tensor([[1.0]]))
ops = [lap, prec, mgp.ScaleWrapperOperator(prec, torch.tensor(2.0)), mgp.NoiseWrapperOperator(prec, torch.tensor(0.1)),
       mgp.SchurComplementOperator(prec, torch.tensor([True, False, True, True, False, True]))]
rhs = torch.rand(n, 2)
for op in ops:
    assert isinstance(op, linear_operator.LinearOperator)
    op.solve(rhs); op.inv_quad_logdet(inv_quad_rhs=rhs, logdet=True); op.logdet(); op.inv_quad(rhs)
prec.diagonalization(); lap.diagonalization("lanczos", 2)
names = [c[0] for c in calls]
assert names.count("solve") == 5 and names.count("iql") == 15 and names.count("diag") == 2, calls
print("REAL_BASE_OK")
'''


def test_cuda_solvers_stay_in_front_of_a_real_linear_operator_base(tmp_path):
    """With ``linear_operator`` importable the operators subclass ITS LinearOperator (gpytorch's isinstance checks hold), and
    ``solve / inv_quad_logdet / logdet / inv_quad / diagonalization`` must still reach manifold_gp_b200.solvers (the CUDA CG /
    SLQ / Lanczos drivers), not the base class's eager loops -- reference call sites precision_matern_operator.py:53,
    schur_complement_operator.py:28, utils/train_model.py:55,67-68."""
    import os
    import subprocess
    import sys
    pkg = tmp_path / "linear_operator"
    pkg.mkdir()
    (pkg / "__init__.py").write_text(_FAKE_LINEAR_OPERATOR)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([str(tmp_path), root, os.environ.get("PYTHONPATH", "")]))
    r = subprocess.run([sys.executable, "-c", _REAL_BASE_SCRIPT], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "REAL_BASE_OK" in r.stdout, r.stdout + r.stderr


def _emulate_pair_walk(t, a, xs_tiles, n, R=128):
    """The lane walk of the paired-row kernel on the CPU: per warp block, step and lane one union entry feeds the lane's OWN
    row and the OTHER row of its pair; lanes 0-3 of a slot output row A, lanes 4-7 row B."""
    qptr, qcol, qsrc, qrow = t["qptr"].long(), t["qcol"].long() & 0xFFFF, t["qsrc"].long().view(-1, 2), t["qrow"].long()
    ntiles = (n + R - 1) // R
    nnzq = t["nnzq"]
    pos = torch.arange(nnzq)
    blk = torch.searchsorted(qptr[:16 * ntiles + 1], pos, right=True) - 1
    lane = (pos - qptr[blk]) & 31
    tile = blk >> 4
    a0 = torch.cat([a, torch.zeros(1, dtype=a.dtype)])
    vm, vo = a0[qsrc[:nnzq, 0]], a0[qsrc[:nnzq, 1]]            # index -1 -> the appended zero
    xr = xs_tiles[tile, qcol[:nnzq]]                             # [nnzq, C]
    pi = (blk & 15) * 4 + (lane >> 3)
    h = (lane >> 2) & 1
    row_mine = tile * R + qrow[tile * R + 2 * pi + h]
    row_other = tile * R + qrow[tile * R + 2 * pi + (1 - h)]
    y = torch.zeros(ntiles * R, xs_tiles.shape[2], dtype=a.dtype)
    y.index_add_(0, row_mine, vm.unsqueeze(1) * xr)
    y.index_add_(0, row_other, vo.unsqueeze(1) * xr)
    return y[:n], dict(lane=lane, blk=blk, col=qcol[:nnzq], real=(qsrc[:nnzq] >= 0).any(1))


def test_pair_streams_reproduce_the_matvec_and_keep_lane_parity():
    from manifold_gp_b200.graph import pair_streams
    g = torch.Generator().manual_seed(3)
    R = 128
    for n, maxlen, with_pos in ((300, 9, True), (128, 40, False), (1000, 5, True), (5, 7, True), (257, 50, True), (640, 30, False)):
        ntiles = (n + R - 1) // R
        rowlen = torch.randint(0, maxlen, (n,), generator=g)
        rowptr = torch.zeros(n + 1, dtype=torch.int32)
        rowptr[1:] = torch.cumsum(rowlen, 0)
        nnz = int(rowptr[-1])
        # columns drawn from a narrow window so that neighbouring rows really share columns; a few duplicates inside a row
        rows = torch.repeat_interleave(torch.arange(n), rowlen)
        lcol = torch.cat([(r % R) // 2 + torch.randperm(64, generator=g)[:int(rowlen[r])] for r in range(n)] + [torch.zeros(0, dtype=torch.long)])
        if nnz > 4:
            lcol[1] = lcol[0]                            # a repeated column inside a row (the reference lists diagonal entries twice)
        lcol = lcol.to(torch.int16)
        a = torch.randn(nnz, generator=g, dtype=torch.float64)
        pos = None
        if with_pos:                                   # a spatial order that differs from the row order inside every tile
            pos = torch.cat([tl * R + torch.randperm(min(R, n - tl * R), generator=g) for tl in range(ntiles)])
        t = pair_streams(rowptr, lcol, n, pos, R)
        assert t["qnzmax"] % 32 == 0 and t["qptr"].numel() == 512 * ((ntiles + 31) // 32) + 4
        xs = torch.randn(ntiles, 256, 4, generator=g, dtype=torch.float64)
        y, info = _emulate_pair_walk(t, a, xs, n)
        ref = torch.zeros(n, 4, dtype=torch.float64)
        ref.index_add_(0, rows, a.unsqueeze(1) * xs[rows // R, lcol.long()])
        assert torch.allclose(y, ref, atol=1e-12)
        # every CSR entry is used exactly once
        used = t["qsrc"].long()
        used = used[used >= 0]
        assert used.numel() == nnz and torch.equal(torch.sort(used).values, torch.arange(nnz))
        # the row table is a permutation of every tile's rows
        assert torch.equal(torch.sort(t["qrow"].long().view(ntiles, R), 1).values, torch.arange(R).expand(ntiles, R))
        # padding columns are valid own rows; lanes 0-3 of a slot hold even columns, lanes 4-7 odd ones, up to the spill
        lane, col, real = info["lane"], info["col"], info["real"]
        nrows_t = torch.clamp(n - (info["blk"] >> 4) * R, max=R)
        assert bool((col[~real] < nrows_t[~real]).all())
        expect = (lane >> 2) & 1
        mism = ((col & 1) != expect) & real
        assert float(mism.sum()) <= 0.5 * float(real.sum())
    # Morton-adjacent rows with identical lists: the union is half the entries
    n = 256
    rowptr = torch.arange(0, 8 * (n + 1), 8, dtype=torch.int32)
    lcol = (torch.arange(n).repeat_interleave(8) // 2 * 2 + torch.arange(8).repeat(n) * 16).remainder(300).to(torch.int16)
    t = pair_streams(rowptr, lcol, n, None, R)
    assert t["q_unions"] == 8 * n // 2


def test_pair_matching_is_a_perfect_matching_that_beats_adjacent_rows():
    from manifold_gp_b200.graph import pair_matching, pair_streams
    g = torch.Generator().manual_seed(4)
    R = 128
    for n in (300, 256, 77):
        ntiles = (n + R - 1) // R
        # rows k and k + 64 of a tile share most of their columns: adjacent pairing finds nothing, the matching everything
        rowlen = torch.full((n,), 12)
        rowptr = torch.zeros(n + 1, dtype=torch.int32)
        rowptr[1:] = torch.cumsum(rowlen, 0)
        rows = torch.repeat_interleave(torch.arange(n), rowlen)
        grp = (rows % R) % 64
        lcol = (grp * 4 + torch.arange(12).repeat(n) % 12 + (torch.rand(rows.numel(), generator=g) < 0.1) * 300).to(torch.int16)
        # make the columns of a row distinct
        lcol = (grp * 12 + torch.arange(12).repeat(n)).to(torch.int16)
        pos = pair_matching(rowptr, lcol, n, R)
        assert pos.numel() == ntiles * R
        assert torch.equal(torch.sort(pos.view(ntiles, R), 1).values, torch.arange(R).expand(ntiles, R))
        t_match = pair_streams(rowptr, lcol, n, pos, R)
        t_adj = pair_streams(rowptr, lcol, n, None, R)
        assert t_match["q_unions"] < t_adj["q_unions"]
        a = torch.randn(int(rowptr[-1]), generator=g, dtype=torch.float64)
        xs = torch.randn(ntiles, 800, 3, generator=g, dtype=torch.float64)
        y, _ = _emulate_pair_walk(t_match, a, xs, n)
        ref = torch.zeros(n, 3, dtype=torch.float64)
        ref.index_add_(0, rows, a.unsqueeze(1) * xs[rows // R, lcol.long()])
        assert torch.allclose(y, ref, atol=1e-12)
    # full tiles: every row k is matched with k + 64 (identical column lists)
    n = 256
    rowptr = torch.arange(0, 12 * (n + 1), 12, dtype=torch.int32)
    rows = torch.repeat_interleave(torch.arange(n), 12)
    lcol = (((rows % R) % 64) * 12 + torch.arange(12).repeat(n)).to(torch.int16)
    pos = pair_matching(rowptr, lcol, n, R).view(2, R)
    for k in range(64):
        assert int(pos[0, k]) // 2 == int(pos[0, k + 64]) // 2


def test_pair_matching_is_chunk_independent():
    """The matching is computed per tile; the chunking of tiles (bounded dense pattern memory) must not change it."""
    from manifold_gp_b200.graph import pair_matching
    g = torch.Generator().manual_seed(7)
    R, n = 128, 700
    rowlen = torch.randint(3, 20, (n,), generator=g)
    rowptr = torch.zeros(n + 1, dtype=torch.int32)
    rowptr[1:] = torch.cumsum(rowlen, 0)
    lcol = torch.cat([(r % R) // 3 + torch.randperm(48, generator=g)[:int(rowlen[r])] for r in range(n)]).to(torch.int16)
    a = pair_matching(rowptr, lcol, n, R, chunk_tiles=2048)
    b = pair_matching(rowptr, lcol, n, R, chunk_tiles=1)
    c = pair_matching(rowptr, lcol, n, R, chunk_tiles=4)
    assert torch.equal(a, b) and torch.equal(a, c)
