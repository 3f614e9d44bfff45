"""sequoia_b200: B200-native (sm_100a) implementation of SEQUOIA's data-parallel hot paths.

Host side mirrors the reference's Python surface (src/tformer_lin.py, src/resnet.py,
pre_processing/kmean_features.py, pre_processing/compute_features_hdf5.py); all arithmetic runs in
hand-written CUDA behind the C ABI in include/sequoia_b200.h.  There is no CPU fallback.
"""
from . import _lib  # noqa: F401  (fails loudly when the shared library has not been built)

__version__ = "0.1.0"
