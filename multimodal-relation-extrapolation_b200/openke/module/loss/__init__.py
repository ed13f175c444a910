from .MarginLoss import MarginLoss

__all__ = ["MarginLoss"]
