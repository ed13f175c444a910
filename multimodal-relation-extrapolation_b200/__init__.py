"""mre_b200 -- B200-native (sm_100a) link-prediction scorer + filtered ranker, Bernoulli negative sampler and
TransE margin-loss step behind the reference's Python API (OpenKE loaders / Tester / Trainer, the paper's
evaluate), calling the C-ABI library build/libmre_b200.so (include/mre_b200.h).

The directory name carries a hyphen (the repository contract); import it as `mre_b200` (the alias module at the
repo root) or with importlib.
"""
from . import _lib  # noqa: F401
from ._lib import MreError, build  # noqa: F401

__all__ = ["_lib", "MreError", "build", "engine", "openke", "paper", "dist"]


def __getattr__(name):
    # torch-dependent submodules load lazily so that `build()` works in a bare interpreter
    if name in ("engine", "openke", "paper", "dist"):
        import importlib
        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(name)
