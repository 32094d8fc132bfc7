"""Numerical knobs of the solvers, mirroring the ``gpytorch.settings`` context managers the reference's callers use
(utils/train_model.py:54,66; utils/test_model.py:11; operators/graph_laplacian_operator.py:133).

If gpytorch is importable its own settings objects are re-exported, so ``with gpytorch.settings.cg_tolerance(1e-2):``
in user code steers the CUDA solvers unchanged.  Otherwise (this image: gpytorch / linear_operator are not installed)
the local stand-ins below provide the same names, defaults and ``.value()`` / ``.on()`` protocol.
"""
from __future__ import annotations

try:  # pragma: no cover - not installed in this image
    from gpytorch.settings import (  # noqa: F401
        cg_tolerance, eval_cg_tolerance, max_cg_iterations, max_cholesky_size, max_lanczos_quadrature_iterations,
        max_root_decomposition_size, num_trace_samples, terminate_cg_by_size, fast_pred_var, _use_eval_tolerance,
    )
    HAVE_GPYTORCH = True
except Exception:  # ImportError or partial install
    HAVE_GPYTORCH = False

    class _value_context:
        _global_value = None

        @classmethod
        def value(cls):
            return cls._global_value

        @classmethod
        def _set_value(cls, value):
            cls._global_value = value

        def __init__(self, value):
            self._orig_value = self.__class__.value()
            self._instance_value = value

        def __enter__(self):
            self.__class__._set_value(self._instance_value)
            return self

        def __exit__(self, *args):
            self.__class__._set_value(self._orig_value)
            return False

    class _feature_flag:
        _default = False
        _state = None

        @classmethod
        def on(cls):
            return cls._default if cls._state is None else cls._state

        @classmethod
        def off(cls):
            return not cls.on()

        @classmethod
        def _set_state(cls, state):
            cls._state = state

        def __init__(self, state=True):
            self.prev = self.__class__._state
            self.state = state

        def __enter__(self):
            self.__class__._set_state(self.state)
            return self

        def __exit__(self, *args):
            self.__class__._set_state(self.prev)
            return False

    class cg_tolerance(_value_context):
        _global_value = 1

    class eval_cg_tolerance(_value_context):
        _global_value = 1e-2

    class max_cg_iterations(_value_context):
        _global_value = 1000

    class max_cholesky_size(_value_context):
        _global_value = 800

    class max_lanczos_quadrature_iterations(_value_context):
        _global_value = 20

    class max_root_decomposition_size(_value_context):
        _global_value = 100

    class num_trace_samples(_value_context):
        _global_value = 10

    class terminate_cg_by_size(_feature_flag):
        _default = False

    class fast_pred_var(_feature_flag):
        _default = False

    class _use_eval_tolerance(_feature_flag):
        _default = False


class cg_check_interval:
    """How many CG iterations the host enqueues between polls of the device-side convergence flag."""
    _global_value = 16

    @classmethod
    def value(cls):
        return cls._global_value


class cg_cuda_graph:
    """Replay the CG iterations of long solves from a captured CUDA graph (one graph = ``cg_check_interval`` iterations).
    The first chunk always runs eagerly, so short solves never pay for a capture."""
    _global_value = True

    @classmethod
    def on(cls):
        return cls._global_value

    def __init__(self, state: bool = True):
        self.state = bool(state)

    def __enter__(self):
        self.prev = cg_cuda_graph._global_value
        cg_cuda_graph._global_value = self.state
        return self

    def __exit__(self, *exc):
        cg_cuda_graph._global_value = self.prev
        return False


class cg_polish:
    """One true-residual correction after a converged fp32 CG solve (``"auto"``, default), always (``True``) or never
    (``False``).

    linear_cg (and the reference's fp32 mBCG with it) stops on the RECURRENCE residual; over ~10^3 fp32 iterations of an
    operator with condition number ~3 * 10^4 the recurrence drifts from b - A x (cfg-C: recurrence 1e-6, true residual 2.3e-4).
    The drift sits in the high modes, which CG removes in a handful of iterations: the polish evaluates r = b - A x once,
    solves A d = r with the same kernels (>= 11 iterations by the published minimum-iteration rule) and returns x + d.
    ``"auto"`` polishes fp32 solves that converged with tolerance <= 1e-4; the iteration counts / tridiagonals reported are
    those of the main run (SLQ is unaffected), ``info["polish"]`` carries the extra iterations and the residuals."""
    _global_value = "auto"

    @classmethod
    def value(cls):
        return cls._global_value

    def __init__(self, state="auto"):
        self.state = state

    def __enter__(self):
        self.prev = cg_polish._global_value
        cg_polish._global_value = self.state
        return self

    def __exit__(self, *exc):
        cg_polish._global_value = self.prev
        return False
