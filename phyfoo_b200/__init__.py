"""phyfoo_b200 -- B200-native drop-in for PhyFoo's site-scoring / free-energy hot path.

All arithmetic runs in phyfoo_b200/libphyfoo.so (CUDA, sm_100a) behind the C ABI of include/phyfoo.h;
this package is the host-side mirror of the reference's model interface.  There is no CPU fallback.
"""
from . import _abi  # noqa: F401
from .lib import IllegalArgumentException, PhyfooError, build  # noqa: F401
