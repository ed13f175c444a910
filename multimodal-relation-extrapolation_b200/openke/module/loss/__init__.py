from .MarginLoss import MarginLoss
from .SigmoidLoss import SigmoidLoss
from .SoftplusLoss import SoftplusLoss

__all__ = ["MarginLoss", "SigmoidLoss", "SoftplusLoss"]
