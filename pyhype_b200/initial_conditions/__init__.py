from .base import InitialCondition
from .box import BoxInitialCondition
from .supersonic_flood import SupersonicFloodInitialCondition

__all__ = ["InitialCondition", "BoxInitialCondition", "SupersonicFloodInitialCondition"]
