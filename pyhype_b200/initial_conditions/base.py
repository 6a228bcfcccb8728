"""pyhype/initial_conditions/base.py:28-31"""
from abc import ABC, abstractmethod


class InitialCondition(ABC):
    @abstractmethod
    def apply_to_block(self, block):
        raise NotImplementedError("Abstract Initial Condition")
