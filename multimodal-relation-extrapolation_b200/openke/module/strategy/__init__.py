from .NegativeSampling import NegativeSampling

__all__ = ["NegativeSampling"]
