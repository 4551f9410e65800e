"""B200-native drop-in for the ``trmf`` package of rofuyu/exp-trmf-nips16
(same exports as the reference's python/trmf/__init__.py:2-6)."""
from .trmf import Model, Metrics, NormalizedTransform
from .trmf import train, rolling_validate, grid_search

__all__ = ["Model", "Metrics", "rolling_validate", "grid_search", "train", "NormalizedTransform"]
