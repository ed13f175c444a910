"""Drop-in mirror of the OpenKE surface the reference's hot path goes through (vendored under /root/reference/OpenKE):
same module layout, class names, constructor arguments, dict-of-arrays batches and metric tuple -- with the native
runtime Base.so replaced by libmre_b200.so and the eager torch scoring ops by its kernels.

    from mre_b200.openke.config import Trainer, Tester
    from mre_b200.openke.module.model import TransE, DistMult, ComplEx
    from mre_b200.openke.module.loss import MarginLoss
    from mre_b200.openke.module.strategy import NegativeSampling
    from mre_b200.openke.data import TrainDataLoader, TestDataLoader
"""
from . import config, data, module  # noqa: F401
