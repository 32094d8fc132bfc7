"""LinearOperator classes of the hot path (same names and constructor signatures as ``manifold_gp.operators``); every
product goes through the CUDA kernels of ``libmgp_b200`` -- see each module for the reference lines it replaces."""
from .noise_wrapper_operator import NoiseWrapperOperator
from .scale_wrapper_operator import ScaleWrapperOperator
from .schur_complement_operator import SchurComplementOperator
from .precision_matern_operator import PrecisionMaternOperator
from .graph_laplacian_operator import GraphLaplacianOperator

__all__ = sorted(name for name, obj in list(globals().items()) if isinstance(obj, type) and name.endswith("Operator"))
