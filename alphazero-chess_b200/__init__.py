"""alphazero-chess_b200: B200-native batched self-play engine (host-side Python mirror over the C ABI)."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libaz_b200.so")
_lib = None


def lib():
    """Loads the CUDA extension; there is no CPU fallback, a missing library is an error."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run __graft_entry__.build() (nvcc, sm_100a) first")
        _lib = ctypes.CDLL(LIB_PATH)
    return _lib
