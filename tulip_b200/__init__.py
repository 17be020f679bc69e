"""tulip_b200 -- B200-native (sm_100a) TULIP Swin forward/backward path.

The product is `tulip_b200/lib/libtulip_b200.so` (hand-written CUDA behind the C ABI in
include/tulip_b200.h) plus this thin host layer mirroring the reference's module API:

    from tulip_b200.model.tulip import tulip_base, tulip_large, TULIP

There is no CPU or PyTorch fallback: importing works anywhere, but every compute call raises
if the library is missing or no CUDA device is present.
"""
from ._lib import TulipLibraryError, lib_path, load_library  # noqa: F401

__all__ = ["TulipLibraryError", "lib_path", "load_library"]
