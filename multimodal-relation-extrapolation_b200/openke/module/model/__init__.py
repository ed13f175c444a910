from .Model import Model
from .TransE import TransE
from .DistMult import DistMult
from .ComplEx import ComplEx
from .SimplE import SimplE
from .TransH import TransH
from .TransD import TransD
from .Analogy import Analogy
from .RotatE import RotatE

__all__ = ["Model", "TransE", "DistMult", "ComplEx", "SimplE", "TransH", "TransD", "Analogy", "RotatE"]
