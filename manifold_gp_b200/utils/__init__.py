from .nearest_neighbors import NearestNeighbors
from .torch_utils import bump_function

__all__ = ["NearestNeighbors", "bump_function"]
