"""argtypes of the remaining C-ABI entry points (filled in as kernels land)."""
import ctypes as C

SIGS = {}
