"""Process-wide precision switch.

"bf16" (default): activations/weights stored in bf16, convolutions and linear layers on the tcgen05
tensor cores with fp32 accumulation in TMEM; statistics, reductions and losses in fp32.
"fp32": everything in fp32 on the CUDA cores (the <= 1e-4 parity mode of BASELINE.json's north_star).
"""
import contextlib

import torch

_PRECISION = "bf16"


def set_precision(p):
    global _PRECISION
    if p not in ("bf16", "fp32"):
        raise ValueError("precision must be 'bf16' or 'fp32'")
    _PRECISION = p


def precision():
    return _PRECISION


def act_dtype():
    return torch.bfloat16 if _PRECISION == "bf16" else torch.float32


@contextlib.contextmanager
def use_precision(p):
    old = _PRECISION
    set_precision(p)
    try:
        yield
    finally:
        set_precision(old)


# Debug aid for tests: run bf16 convolutions on the CUDA-core kernel instead of the tcgen05 kernel, to
# separate "bf16 rounding noise" from "tensor-core path bug" when a bf16 parity number looks large.
FORCE_SIMT = False
