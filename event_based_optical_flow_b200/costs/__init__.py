"""Cost plugin registry, keyed by `.name` exactly like the reference (src/costs/__init__.py:23-38)."""
from .base import CostBase
from .contrast import (GradientMagnitude, ImageVariance, MultiFocalNormalizedGradientMagnitude,
                       MultiFocalNormalizedImageVariance, NormalizedGradientMagnitude, NormalizedImageVariance)
from .total_variation import TotalVariation


def inheritors(klass):
    subclasses = set()
    work = [klass]
    while work:
        parent = work.pop()
        for child in parent.__subclasses__():
            if child not in subclasses:
                subclasses.add(child)
                work.append(child)
    return subclasses


functions = {k.name: k for k in inheritors(CostBase) if hasattr(k, "name")}

from .hybrid import HybridCost  # noqa: E402  (needs `functions`)
