from .Model import Model
from .TransE import TransE
from .DistMult import DistMult
from .ComplEx import ComplEx

__all__ = ["Model", "TransE", "DistMult", "ComplEx"]
