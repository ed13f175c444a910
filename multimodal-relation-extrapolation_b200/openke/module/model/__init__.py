from .Model import Model
from .TransE import TransE
from .DistMult import DistMult
from .ComplEx import ComplEx
from .SimplE import SimplE

__all__ = ["Model", "TransE", "DistMult", "ComplEx", "SimplE"]
