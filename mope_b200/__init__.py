"""mope_b200 -- B200-native (sm_100a) implementation of mope's ontogenetic-phylogeny log-likelihood.

Python host + hand-written CUDA kernels behind the C ABI in include/mope_b200.h; PyTorch owns device buffers.
"""
__version__ = "0.1.0"

from .tables import TransitionTables, load as load_tables  # noqa: F401
from .inference import Inference, GpuPool  # noqa: F401
from .engine import Engine, OutOfTableError  # noqa: F401
