"""manifold_gp_b200 -- B200-native (sm_100a) implementation of the data-parallel hot path of IMGP
(nash169/manifold-gp): kNN graph build -> graph-Laplacian operator -> Matern precision operator -> CG / Lanczos,
behind the reference's own operator / kernel / model surface.

Importing the package loads libmgp_b200.so (built in-tree by ``manifold_gp_b200/csrc/build.py``); there is no
PyTorch or CPU fallback for any operation on the path.
"""
from . import _lib  # noqa: F401  (fails loudly if the CUDA library is missing)
from . import settings  # noqa: F401
from .operators import (  # noqa: F401
    GraphLaplacianOperator,
    NoiseWrapperOperator,
    PrecisionMaternOperator,
    ScaleWrapperOperator,
    SchurComplementOperator,
)
from .utils import NearestNeighbors, bump_function  # noqa: F401
from .kernels import RiemannKernel, RiemannMaternKernel  # noqa: F401
from .models import RiemannGP  # noqa: F401
from ._compat.gp import GaussianLikelihood, ScaleKernel  # noqa: F401  (stand-ins; use gpytorch's when it is installed)

__version__ = "0.1.0"
