"""``LinearOperator`` base used by manifold_gp_b200.operators.

The reference's operators subclass ``linear_operator.LinearOperator`` (manifold_gp/operators/*.py).  When that
package is importable it is used as the base class unchanged, so the operators drop into GPyTorch exactly like the
reference's.  In this image neither gpytorch nor linear_operator is installed, so a minimal protocol base with the
same method names is provided: ``_matmul / _size / _transpose_nonbatch / _diagonal`` are what subclasses implement
and ``matmul / solve / inv_quad_logdet / diagonalization / to_dense / diagonal / T`` are what callers use
(utils/train_model.py:55,67-68; schur_complement_operator.py:28; graph_laplacian_operator.py:135;
test/_test_functions.py).  The solver entry points dispatch exactly like linear_operator's: dense Cholesky / eigh
when the size is <= ``settings.max_cholesky_size``, otherwise the CUDA mBCG / Lanczos drivers in
``manifold_gp_b200.solvers``.
"""
from __future__ import annotations

import torch

try:
    from linear_operator import LinearOperator as _RealLinearOperator
    HAVE_LINEAR_OPERATOR = True
except Exception:
    _RealLinearOperator = None
    HAVE_LINEAR_OPERATOR = False


def _flatten_tensors(obj, out):
    if torch.is_tensor(obj):
        out.append(obj)
    elif isinstance(obj, LocalLinearOperator):
        obj._collect_tensors(out)
    elif isinstance(obj, (list, tuple)):
        for o in obj:
            _flatten_tensors(o, out)


class LocalLinearOperator:
    """Minimal stand-in for ``linear_operator.LinearOperator`` (non-batch, square, real)."""

    def __init__(self, *args, **kwargs):
        self._args = args
        self._kwargs = kwargs

    # ---- to be provided by subclasses --------------------------------------------------------------------------
    def _matmul(self, rhs):
        raise NotImplementedError

    def _size(self):
        raise NotImplementedError

    def _transpose_nonbatch(self):
        raise NotImplementedError

    def _diagonal(self):
        return self.to_dense().diagonal()

    # ---- representation -------------------------------------------------------------------------------------------
    def _collect_tensors(self, out):
        for a in self._args:
            _flatten_tensors(a, out)
        for a in self._kwargs.values():
            _flatten_tensors(a, out)

    def representation(self):
        out = []
        self._collect_tensors(out)
        return tuple(out)

    def _float_tensors(self):
        return [t for t in self.representation() if t.is_floating_point()]

    @property
    def dtype(self):
        ts = self._float_tensors()
        return ts[0].dtype if ts else torch.get_default_dtype()

    @property
    def device(self):
        ts = self.representation()
        return ts[0].device if ts else torch.device("cpu")

    @property
    def requires_grad(self):
        return any(t.requires_grad for t in self._float_tensors())

    # ---- shape ----------------------------------------------------------------------------------------------------
    @property
    def shape(self):
        return torch.Size([int(s) for s in self._size()])

    def size(self, dim=None):
        return self.shape if dim is None else self.shape[dim]

    def dim(self):
        return 2

    ndimension = dim

    def numel(self):
        return self.shape[0] * self.shape[1]

    @property
    def is_square(self):
        return self.shape[0] == self.shape[1]

    @property
    def matrix_shape(self):
        return self.shape

    @property
    def batch_shape(self):
        return torch.Size([])

    # ---- algebra --------------------------------------------------------------------------------------------------
    def matmul(self, other):
        if not torch.is_tensor(other):
            raise TypeError("LocalLinearOperator.matmul expects a tensor right-hand side")
        return self._matmul(other)

    __matmul__ = matmul

    def t(self):
        return self._transpose_nonbatch()

    @property
    def T(self):
        return self._transpose_nonbatch()

    @property
    def mT(self):
        return self._transpose_nonbatch()

    def transpose(self, d1, d2):
        nd = 2
        d1, d2 = d1 % nd, d2 % nd
        return self if d1 == d2 else self._transpose_nonbatch()

    def diagonal(self, offset=0, dim1=-2, dim2=-1):
        if offset != 0:
            raise NotImplementedError("only the main diagonal is available")
        return self._diagonal()

    def to_dense(self):
        n = self.shape[1]
        return self._matmul(torch.eye(n, dtype=self.dtype, device=self.device))

    def evaluate(self):  # old gpytorch spelling
        return self.to_dense()

    def detach(self):
        return self



class CudaSolverMixin:
    """``solve / inv_quad_logdet / inv_quad / logdet / diagonalization`` routed to the CUDA drivers of ``manifold_gp_b200.solvers``.

    Placed BEFORE the base class in the MRO of every operator of this package, so the routing also holds when the real
    ``linear_operator.LinearOperator`` is the base: without it, ``op.solve`` / ``op.inv_quad_logdet`` (utils/train_model.py:55,67-68,
    precision_matern_operator.py:53, schur_complement_operator.py:28) would run linear_operator's eager ``linear_cg`` -- ten small
    launches and one host sync per iteration -- over ``_matmul``, bypassing the fused CG / Lanczos / SLQ kernels.  Dispatch rules
    (dense Cholesky / eigh when size <= ``settings.max_cholesky_size``) are linear_operator's own."""

    def solve(self, right_tensor, left_tensor=None):
        from .. import solvers
        out = solvers.solve(self, right_tensor)
        return out if left_tensor is None else left_tensor @ out

    def inv_quad_logdet(self, inv_quad_rhs=None, logdet=False, reduce_inv_quad=True):
        from .. import solvers
        return solvers.inv_quad_logdet(self, inv_quad_rhs=inv_quad_rhs, logdet=logdet, reduce_inv_quad=reduce_inv_quad)

    def inv_quad(self, inv_quad_rhs, reduce_inv_quad=True):
        return self.inv_quad_logdet(inv_quad_rhs=inv_quad_rhs, logdet=False, reduce_inv_quad=reduce_inv_quad)[0]

    def logdet(self):
        return self.inv_quad_logdet(inv_quad_rhs=None, logdet=True)[1]

    def diagonalization(self, method=None):
        from .. import solvers
        return solvers.diagonalization(self, method=method)


class LinearOperator(CudaSolverMixin, _RealLinearOperator if HAVE_LINEAR_OPERATOR else LocalLinearOperator):
    """Base class of every operator in ``manifold_gp_b200.operators``: the real ``linear_operator.LinearOperator`` when that
    package is importable (``isinstance`` checks of gpytorch keep working), the local stand-in otherwise -- in both cases with
    the solver entry points of ``CudaSolverMixin`` in front (tests/test_host_logic.py runs the real-base branch against a
    stand-in ``linear_operator`` module)."""


class DenseEigenvectors:
    """What ``diagonalization`` returns for eigenvectors: the reference calls ``.to_dense()`` on it
    (graph_laplacian_operator.py:139)."""

    def __init__(self, tensor):
        self.tensor = tensor

    def to_dense(self):
        return self.tensor

    @property
    def shape(self):
        return self.tensor.shape

    def __getitem__(self, item):
        return DenseEigenvectors(self.tensor[item])
