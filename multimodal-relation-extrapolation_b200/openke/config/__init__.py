from .Trainer import Trainer
from .Tester import Tester

__all__ = ["Trainer", "Tester"]
