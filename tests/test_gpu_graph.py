"""GPU parity: kNN search, symmetrise/coalesce, CSR build -- through the C ABI, against the oracle."""
import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _mgp():
    import manifold_gp_b200 as mgp
    return mgp


def _check_knn(x, q, k, self_search):
    mgp = _mgp()
    knn = mgp.NearestNeighbors(x.to(DEV))
    d, i = knn.search(q.to(DEV), k)
    d, i = d.cpu(), i.cpu()
    od, oi = oracle.knn_search(x, q, k)
    assert d.shape == (q.shape[0], k) and i.dtype == torch.int64
    assert bool((d[:, 1:] >= d[:, :-1]).all()), "distances not ascending"
    # bit-exact index lists against the fp32 oracle (same arithmetic: mul-then-add in ascending d, ties by index)
    exact = (i == oi).all(1).float().mean().item()
    assert exact == 1.0, f"only {exact:.6f} of rows match the oracle exactly"
    assert torch.equal(d, od)
    # set equality against fp64 ground truth up to near-ties
    d64, i64 = oracle.knn_search_exact(x, q, k)
    bad = ~(i.sort(1).values == i64.sort(1).values).all(1)
    if bad.any():
        # a differing row must be a near-tie at the k-th distance in fp64
        gap = (d64[bad, -1] - d[bad, -1].double()).abs() / d64[bad, -1].clamp_min(1e-30)
        assert float(gap.max()) < 1e-5
    if self_search:
        assert torch.equal(i[:, 0], torch.arange(q.shape[0]))


def test_knn_torus_d3():
    x = oracle.datasets.torus(20000, seed=2)
    _check_knn(x, x, 32, True)


def test_knn_ragged_sizes_and_k():
    x = oracle.datasets.torus(5003, seed=4)
    q = oracle.datasets.torus(131, seed=9)
    _check_knn(x, q, 1, False)
    _check_knn(x, q, 10, False)
    _check_knn(x, x[:777], 64, False)


def test_knn_multi_chunk_dims():
    g = torch.Generator().manual_seed(0)
    for d in (2, 17, 40):
        x = torch.randn(3001, d, generator=g)
        _check_knn(x, x, 9, True)


def test_knn_rmnist_shape_small():
    x = oracle.datasets.rmnist_shape(n=2048, d=784, prototypes=8, seed=1)
    _check_knn(x, x, 10, True)


def test_knn_fewer_points_than_k():
    mgp = _mgp()
    x = torch.tensor([[0.0, 0.0], [1.0, 0.0], [0.0, 2.0]])
    d, i = mgp.NearestNeighbors(x.to(DEV)).search(x.to(DEV), 5)
    assert bool(torch.isinf(d[:, 3:]).all()) and bool((i[:, 3:] == -1).all())
    od, oi = oracle.knn_search(x, x, 5)
    assert torch.equal(i.cpu()[:, :3], oi[:, :3])


def test_graph_matches_oracle_and_golden(dumbbell, golden_k10):
    mgp = _mgp()
    x = dumbbell["train_x"]
    idx, val = mgp.NearestNeighbors(x.to(DEV)).graph(10)
    assert torch.equal(idx.cpu(), torch.from_numpy(golden_k10["idx"]).long())
    assert torch.equal(val.cpu(), torch.from_numpy(golden_k10["val"]))
    # sortedness / uniqueness / upper-triangularity (size-independent properties)
    key = idx[0] * x.shape[0] + idx[1]
    assert bool((idx[0] < idx[1]).all()) and bool((key[1:] > key[:-1]).all())


def test_graph_large_properties():
    mgp = _mgp()
    x = oracle.datasets.torus(200000, seed=0)
    k = 32
    idx, val = mgp.NearestNeighbors(x.to(DEV)).graph(k)
    n = x.shape[0]
    key = idx[0] * n + idx[1]
    assert bool((idx[0] < idx[1]).all()) and bool((key[1:] > key[:-1]).all())
    m = idx.shape[1]
    assert n * (k - 1) / 2 <= m <= n * (k - 1)
    # every node keeps at least its k-1 out-edges
    deg = torch.bincount(idx.reshape(-1), minlength=n)
    assert int(deg.min()) >= k - 1
    # values are the squared distances of the endpoints
    xd = x.to(DEV)
    d2 = (xd[idx[0]] - xd[idx[1]]).square().sum(1)
    assert torch.allclose(val, d2, rtol=1e-5, atol=1e-9)


def test_symmetrize_with_duplicate_points():
    mgp = _mgp()
    x = torch.tensor([[0.0, 0.0], [0.0, 0.0], [1.0, 0.0], [2.0, 0.0], [3.0, 0.5], [3.0, 0.5], [3.0, 0.5]])
    idx, val = mgp.NearestNeighbors(x.to(DEV)).graph(3)
    oidx, oval = oracle.knn_graph(x, 3)
    assert torch.equal(idx.cpu(), oidx) and torch.allclose(val.cpu(), oval)


def test_csr_structure_against_scipy():
    import scipy.sparse as sp
    from manifold_gp_b200 import graph
    x = oracle.datasets.torus(30011, seed=6)
    oidx, oval = oracle.knn_graph(x, 12)
    n = x.shape[0]
    st = graph.GraphStructure(oidx.to(DEV), n)
    m = oidx.shape[1]
    rows = np.concatenate([oidx[0].numpy(), oidx[1].numpy()])
    cols = np.concatenate([oidx[1].numpy(), oidx[0].numpy()])
    eids = np.concatenate([np.arange(m), np.arange(m)])
    order = np.lexsort((cols, rows))
    rowptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(rowptr, rows + 1, 1)
    rowptr = np.cumsum(rowptr)
    assert np.array_equal(st.rowptr.cpu().numpy(), rowptr)
    # entries inside a row may be in any order (the tile build orders them by shared-memory bank parity): compare the
    # (row, col, eid) triples as sets, row by row
    got_rows = np.repeat(np.arange(n), np.diff(st.rowptr.cpu().numpy()))
    got = np.stack([got_rows, st.col.cpu().numpy(), st.eid.cpu().numpy()], 1)
    ref = np.stack([rows[order], cols[order], eids[order]], 1)
    got = got[np.lexsort((got[:, 1], got[:, 0]))]
    assert np.array_equal(got, ref)
    d2 = st.d2csr(oval.to(DEV))
    assert torch.equal(d2.cpu(), oval[st.eid.cpu().long()])
    up = st.upper_pos().cpu()
    assert torch.equal(st.eid.cpu()[up].long(), torch.arange(m))
