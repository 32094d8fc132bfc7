from .nearest_neighbors import NearestNeighbors
from .torch_utils import bump_function
from .train_model import manifold_informed_train
from .test_model import test_model

__all__ = ["NearestNeighbors", "bump_function", "manifold_informed_train", "test_model"]
