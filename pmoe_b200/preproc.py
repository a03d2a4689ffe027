"""GPU input pipeline: the reference dataset's eval-mode frame transform as one C-ABI call.

Reference: `CarlaSegPred.__getitem__` (PMoE/model/data_loader.py:245-308): `Crop(self.crop)` (augmenter.py:43-49, rows
[top:-bottom]) -> `transforms.Resize(self.resize)` on a PIL image -> `transforms.ToTensor()` -> `torch.stack(imgs)`.
`Resize` on a PIL image is Pillow's `Image.resize(..., BILINEAR)`: an antialiased separable triangle filter evaluated in
22-bit fixed point, horizontal pass first with a uint8 intermediate (Pillow `src/libImaging/Resample.c`, third-party:
restated here from its published source; the image's Pillow is 12.2). This module builds Pillow's coefficient tables on the
host exactly as `precompute_coeffs` / `normalize_coeffs_8bpc` do (same double-precision operations in the same order) and
hands them to `pmoe_preprocess_frames`; the device arithmetic is integer, so the result is bit-identical to the reference
transform (tests/test_preproc.py against fixtures made with the real torchvision/Pillow calls).

PNG decoding and the imgaug training-time augmentations stay on the host (out of scope this round).
"""
import ctypes as C
import math

import torch

from ._lib import lib, check, stream_ptr

PRECISION_BITS = 32 - 8 - 2


def _bilinear(x):
    if x < 0.0:
        x = -x
    if x < 1.0:
        return 1.0 - x
    return 0.0


def pillow_bilinear_coeffs(in_size, out_size):
    """precompute_coeffs(in0=0, in1=in_size) + normalize_coeffs_8bpc of Pillow's Resample.c for the BILINEAR filter
    (support 1.0). Returns (bounds int32 (out, 2) = [first input index, count], coefficients int32 (out, ksize), ksize)."""
    scale = filterscale = float(in_size) / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = torch.zeros(out_size, 2, dtype=torch.int32)
    coef = torch.zeros(out_size, ksize, dtype=torch.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        ww = 0.0
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        k = [0.0] * ksize
        for x in range(xmax):
            w = _bilinear((x + xmin - center + 0.5) * ss)
            k[x] = w
            ww += w
        for x in range(xmax):
            if ww != 0.0:
                k[x] /= ww
        bounds[xx, 0], bounds[xx, 1] = xmin, xmax
        for x in range(ksize):
            v = k[x] * (1 << PRECISION_BITS)
            coef[xx, x] = int(-0.5 + v) if k[x] < 0 else int(0.5 + v)
    return bounds, coef, ksize


class FramePreprocessor:
    """crop=(top, bottom), resize=(H, W): the `crop` / `resize` entries of conf/stage_*.yaml (stage_2.yaml:41-48).
    __call__(frames): uint8 (N, Hs, Ws, 3) RGB tensor (CUDA, or pinned host memory: it is uploaded on the current stream)
    -> float32 (N, 3, H, W) on the GPU; with frames of shape (B, T, Hs, Ws, 3) -> (B, T, 3, H, W), the layout
    `torch.stack(imgs)` + the DataLoader's collate produce for the models' `forward(images, ...)`."""

    def __init__(self, crop=(125, 90), resize=(224, 224), device="cuda"):
        self.top, self.bottom = int(crop[0]), int(crop[1])
        self.oh, self.ow = int(resize[0]), int(resize[1])
        self.device = torch.device(device)
        self._tables = {}
        self.lut = (torch.arange(256, dtype=torch.float32) / 255).to(self.device)  # ToTensor: float(v) / 255 (IEEE division)

    def _table(self, hs, ws):
        key = (hs, ws)
        if key not in self._tables:
            hc = hs - self.top - self.bottom
            if hc < 1:
                raise ValueError("crop (%d, %d) leaves no rows of a %d-row frame" % (self.top, self.bottom, hs))
            hb, hk, hks = pillow_bilinear_coeffs(ws, self.ow)
            vb, vk, vks = pillow_bilinear_coeffs(hc, self.oh)
            self._tables[key] = (hc, hb.to(self.device), hk.to(self.device), hks, vb.to(self.device), vk.to(self.device), vks)
        return self._tables[key]

    def __call__(self, frames, out=None, want_u8=False):
        lead = frames.shape[:-3]
        hs, ws, ch = frames.shape[-3:]
        if frames.dtype != torch.uint8 or ch != 3:
            raise TypeError("frames must be uint8 (..., H, W, 3) RGB")
        src = frames.reshape(-1, hs, ws, 3)
        if not src.is_cuda:
            src = src.to(self.device, non_blocking=True)
        src = src.contiguous()
        n = src.shape[0]
        hc, hb, hk, hks, vb, vk, vks = self._table(hs, ws)
        tmp = torch.empty(n, hc, self.ow, 3, dtype=torch.uint8, device=self.device)
        if out is None:
            out = torch.empty(n, 3, self.oh, self.ow, dtype=torch.float32, device=self.device)
        o = out.view(n, 3, self.oh, self.ow)
        u8 = torch.empty(n, self.oh, self.ow, 3, dtype=torch.uint8, device=self.device) if want_u8 else None
        check(lib().pmoe_preprocess_frames(
            src.data_ptr(), n, hs, ws, self.top, self.bottom, self.oh, self.ow, hb.data_ptr(), hk.data_ptr(), hks, vb.data_ptr(),
            vk.data_ptr(), vks, self.lut.data_ptr(), tmp.data_ptr(), o.data_ptr(), o.stride(0), o.stride(1), o.stride(2),
            o.stride(3), None if u8 is None else C.c_void_p(u8.data_ptr()), stream_ptr()), "preprocess_frames")
        res = out.view(*lead, 3, self.oh, self.ow)
        return (res, u8.view(*lead, self.oh, self.ow, 3)) if want_u8 else res
