from .base import InitialCondition
from .supersonic_flood import SupersonicFloodInitialCondition

__all__ = ["InitialCondition", "SupersonicFloodInitialCondition"]
