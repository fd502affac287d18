"""winnowmap_b200: B200-native (sm_100a) seed-chain-align path of Winnowmap v2.03.

The product is the C-ABI shared library (include/winnowmap_b200.h) built from
winnowmap_b200/csrc; this package is the thin Python binding used by tests and bench.py.
There is no CPU fallback: importing works anywhere, computing requires a CUDA device and
the built extension (python -m winnowmap_b200.build)."""
from ._lib import lib, lib_path  # noqa: F401
from . import kernels  # noqa: F401

__version__ = "0.1.0"
