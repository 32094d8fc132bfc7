"""CPU oracle for the IMGP hot path (kNN graph -> graph Laplacian -> Matern precision -> CG/Lanczos).

TEST INFRASTRUCTURE ONLY.  Nothing under ``manifold_gp_b200/`` imports this package.  The only
callers are ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` -- always as the checker or the timed CPU baseline, never as the product path.

It is a pure-torch (CPU) restatement of the reference's algorithm, one function per reference
function, each citing the reference file:line it follows (paths relative to /root/reference).

Parity status (see DESIGN.md "Oracle"):

* Laplacian / precision / wrappers / Schur / out-of-sample / spectral features: PINNED.  The
  restatement is checked against the reference's own dense oracle ``test/_dense_operators.py``
  (pure torch, importable by file path) on the reference's own fixture ``dumbbell.msh``;
  the outputs of that dense oracle are committed as ``tests/golden/dumbbell_*.npz`` together
  with the generating script ``tests/golden/make_golden.py``.
* kNN search, coalesce, CG (mBCG), Lanczos, SLQ log-det: PARITY UNPINNED against the reference.
  These live in third-party packages (faiss, torch_sparse, linear_operator; un-pinned in
  ``setup.py:26-32``, absent from /root/reference and from this image) and the reference's tests
  never pin them.  They are restated from the published algorithms and pinned
  instead against mathematical ground truth: fp64 brute force (kNN), dense ``torch.linalg.solve``
  / ``eigh`` / ``logdet`` on the assembled matrices (CG / Lanczos / SLQ).
"""

from .knn_graph import knn_search, knn_search_exact, knn_graph, symmetrize_coalesce  # noqa: F401
from .operators import (  # noqa: F401
    LaplacianOracle,
    precision_matmul,
    scale_matmul,
    noise_matmul,
    schur_matmul,
    dense_from_matmul,
)
from .solvers import (  # noqa: F401
    linear_cg,
    lanczos_tridiag,
    lanczos_tridiag_to_diag,
    lanczos_diagonalization,
    inv_quad_logdet,
)
from .kernel import bump_function, spectral_density, eval_eigenpairs, features, low_rank_posterior  # noqa: F401
from . import datasets  # noqa: F401
