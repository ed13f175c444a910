from .TrainDataLoader import TrainDataLoader
from .TestDataLoader import TestDataLoader

__all__ = ["TrainDataLoader", "TestDataLoader"]
