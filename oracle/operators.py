"""Oracle: graph Laplacian, Matern precision and the Scale / Noise / Schur wrappers.  TEST INFRASTRUCTURE ONLY.

Pure-torch CPU restatement of ``manifold_gp/operators/*.py``.  ``torch_sparse.spmm`` is restated as the
ATen sequence it lowers to (index_select -> mul -> scatter_add).  PINNED against the
reference's dense oracle ``test/_dense_operators.py`` through ``tests/golden/dumbbell_*.npz``.

Unlike the reference (graph_laplacian_operator.py:63,67,86,97 allocate in the *default* dtype) every
buffer follows the dtype of the edge values, so fp64 runs do not silently mix precisions.
"""

from __future__ import annotations

import torch


def _spmm(index: torch.Tensor, value: torch.Tensor, m: int, matrix: torch.Tensor) -> torch.Tensor:
    """``torch_sparse.spmm(index, value, m, n, matrix)``: out[row] += value * matrix[col]."""
    row, col = index[0], index[1]
    out = matrix.index_select(0, col) * value.unsqueeze(-1)
    res = torch.zeros(m, matrix.shape[1], dtype=matrix.dtype)
    return res.index_add_(0, row, out)


class LaplacianOracle:
    """``GraphLaplacianOperator`` (graph_laplacian_operator.py:25-157) on CPU tensors.

    ``val``  -- squared kNN distances of the M upper-triangular edges (the reference calls it ``x``),
    ``idx``  -- ``[2, M]`` int64, ``row < col``,
    ``eps``  -- graph bandwidth (python float or 0-d/[1,1] tensor; may require grad).
    """

    def __init__(self, val, idx, n, eps, normalization="randomwalk", self_loops=True, transposed=False):
        self.val = val
        self.idx = idx
        self.n = int(n)
        self.eps = eps if torch.is_tensor(eps) else torch.tensor(float(eps), dtype=val.dtype)
        self.eps = self.eps.reshape(()).to(val.dtype)
        self.normalization = normalization
        self.self_loops = self_loops
        self.transposed = transposed
        self._cache = {}

    def _memo(self, name, fn):
        if name not in self._cache:
            self._cache[name] = fn()
        return self._cache[name]

    def _scatter2(self, base, v):
        return base.index_add(0, self.idx[0], v).index_add(0, self.idx[1], v)

    @property
    def adjacency_unnorm_mat(self):  # :52-56   W = exp(-d^2 / (4 eps^2))
        return self._memo("W", lambda: self.val.div(-4 * self.eps.square()).exp())

    @property
    def degree_unnorm_mat(self):  # :58-69   D~ = [1] + sum_j W_ij
        def f():
            base = torch.ones(self.n, dtype=self.val.dtype) if self.self_loops else torch.zeros(self.n, dtype=self.val.dtype)
            return self._scatter2(base, self.adjacency_unnorm_mat)
        return self._memo("Dt", f)

    @property
    def adjacency_mat(self):  # :71-75   A~ = W / (D~_i D~_j)
        return self._memo("A", lambda: self.adjacency_unnorm_mat.div(
            self.degree_unnorm_mat[self.idx[0]] * self.degree_unnorm_mat[self.idx[1]]))

    @property
    def degree_mat(self):  # :77-88   D = [D~^-2] + sum_j A~_ij
        def f():
            base = self.degree_unnorm_mat.pow(-2) if self.self_loops else torch.zeros(self.n, dtype=self.val.dtype)
            return self._scatter2(base, self.adjacency_mat)
        return self._memo("D", f)

    @property
    def laplacian_diag(self):  # :90-97
        def f():
            if self.self_loops:
                return (1 - self.degree_unnorm_mat.pow(-2) * self.degree_mat.pow(-1)).div(self.eps.square())
            return torch.ones(self.n, dtype=self.val.dtype).div(self.eps.square())
        return self._memo("diag", f)

    @property
    def laplacian_triu(self):  # :102-106
        def f():
            ds = self.degree_mat.sqrt()
            return self.adjacency_mat.div(ds[self.idx[0]] * ds[self.idx[1]]).div(self.eps.square())
        return self._memo("triu", f)

    def matmul(self, rhs: torch.Tensor) -> torch.Tensor:  # _matmul :108-124
        squeeze = rhs.dim() == 1
        if squeeze:
            rhs = rhs.unsqueeze(-1)
        if self.normalization == "randomwalk":
            sq = self.degree_mat.pow(0.5).view(-1, 1)
            vec = rhs.contiguous().div(sq) if self.transposed else rhs.contiguous() * sq
        else:
            vec = rhs.contiguous()
        out = vec * self.laplacian_diag.view(-1, 1)
        out = out - _spmm(self.idx, self.laplacian_triu, self.n, vec)
        out = out - _spmm(torch.stack((self.idx[1], self.idx[0]), 0), self.laplacian_triu, self.n, vec)
        if self.normalization == "randomwalk":
            out = out * (self.degree_mat.pow(0.5).view(-1, 1) if self.transposed else self.degree_mat.pow(-0.5).view(-1, 1))
        return out.squeeze(-1) if squeeze else out

    def transpose(self):  # _transpose_nonbatch :129-130
        if self.normalization == "randomwalk":
            return LaplacianOracle(self.val, self.idx, self.n, self.eps, self.normalization, self.self_loops, not self.transposed)
        return self

    def symmetric_twin(self):
        return LaplacianOracle(self.val, self.idx, self.n, self.eps, "symmetric", self.self_loops)

    def out_of_sample(self, phi: torch.Tensor, edge_value: torch.Tensor, edge_idx: torch.Tensor) -> torch.Tensor:
        """Nystrom extension of eigenvectors ``phi[N,m]`` to Q new points (:146-157)."""
        out = edge_value.div(-4 * self.eps.square()).exp()
        degree_test = out.sum(dim=1)
        out = out / (self.degree_unnorm_mat[edge_idx] * degree_test.view(-1, 1))
        if self.normalization == "symmetric":
            out = out / (self.degree_mat.sqrt()[edge_idx] * out.sum(dim=1).sqrt().view(-1, 1))
        elif self.normalization == "randomwalk":
            out = out / out.sum(dim=1).view(-1, 1)
        return out.unsqueeze(-1).mul(phi[edge_idx]).sum(dim=1)

    def dense(self) -> torch.Tensor:
        return self.matmul(torch.eye(self.n, dtype=self.val.dtype))


def precision_matmul(lap: LaplacianOracle, nu: int, lengthscale, rhs: torch.Tensor) -> torch.Tensor:
    """``PrecisionMaternOperator._matmul`` (precision_matern_operator.py:26-37)."""
    squeeze = rhs.dim() == 1
    out = rhs.unsqueeze(-1) if squeeze else rhs
    ls = lengthscale if torch.is_tensor(lengthscale) else torch.tensor(float(lengthscale), dtype=rhs.dtype)
    diag = ls.reshape(()).to(rhs.dtype).square() / (2 * nu)
    for _ in range(nu):
        out = out + diag * lap.matmul(out)
        out = out / diag
    if lap.normalization == "randomwalk":
        out = out * lap.degree_mat.view(-1, 1)
    return out.squeeze(-1) if squeeze else out


def scale_matmul(inner, scale, rhs, inverse_scale=False):
    """``ScaleWrapperOperator._matmul`` (scale_wrapper_operator.py:27-28)."""
    return inner(rhs) / scale if inverse_scale else inner(rhs) * scale


def noise_matmul(inner, noise, rhs):
    """``NoiseWrapperOperator._matmul`` (noise_wrapper_operator.py:21-22): Q - s Q^2 + s^2 Q^3."""
    return inner(rhs - noise * inner(rhs - noise * inner(rhs)))


def schur_matmul(inner, mask: torch.Tensor, rhs: torch.Tensor, solve=None) -> torch.Tensor:
    """``SchurComplementOperator._matmul`` (schur_complement_operator.py:26-30).

    ``inner`` is the full-size matvec, ``mask`` the labelled rows.  The inner solve on the unlabelled block
    (reference: ``MaskedLinearOperator(...).solve`` = Cholesky or CG) is a dense solve here unless ``solve`` given.
    """
    n = mask.shape[0]
    squeeze = rhs.dim() == 1
    r = rhs.unsqueeze(-1) if squeeze else rhs
    full = torch.zeros(n, r.shape[1], dtype=r.dtype)
    full[mask] = r
    tmp = inner(full)
    if solve is None:
        def zz(v):
            f = torch.zeros(n, v.shape[1], dtype=v.dtype)
            f[~mask] = v
            return inner(f)[~mask]
        qzz = zz(torch.eye(int((~mask).sum()), dtype=r.dtype))
        out = torch.linalg.solve(qzz, tmp[~mask])
    else:
        out = solve(tmp[~mask])
    f = torch.zeros(n, r.shape[1], dtype=r.dtype)
    f[~mask] = out
    res = tmp[mask] - inner(f)[mask]
    return res.squeeze(-1) if squeeze else res


def dense_from_matmul(matmul, n: int, dtype=torch.float64) -> torch.Tensor:
    """``to_dense()`` as linear_operator does it for matrix-free operators: ``_matmul(eye(N))``."""
    return matmul(torch.eye(n, dtype=dtype))
