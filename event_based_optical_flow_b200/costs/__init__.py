"""Cost plugins on the CUDA operators.  `functions` maps a plugin's `name` to its class, the table the reference's solver
indexes with the YAML's `cost:` string (src/costs/__init__.py:35); every subclass of `CostBase` that defines `name`
registers itself on import."""
from .base import REGISTRY as functions
from .base import CostBase
from .contrast import (GradientMagnitude, ImageVariance, MultiFocalNormalizedGradientMagnitude,
                       MultiFocalNormalizedImageVariance, NormalizedGradientMagnitude, NormalizedImageVariance)
from .hybrid import HybridCost
from .total_variation import TotalVariation

__all__ = ["functions", "CostBase", "HybridCost", "TotalVariation", "ImageVariance", "GradientMagnitude", "NormalizedImageVariance",
           "NormalizedGradientMagnitude", "MultiFocalNormalizedImageVariance", "MultiFocalNormalizedGradientMagnitude"]
