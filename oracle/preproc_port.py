"""TEST INFRASTRUCTURE ONLY — CPU restatement (numpy, integer arithmetic) of the reference dataset's eval-mode frame
transform: Crop (PMoE/model/augmenter.py:43-49) -> torchvision Resize on a PIL image (PMoE/model/data_loader.py:275-281)
-> ToTensor. The resize is Pillow's ImagingResample (third-party, not vendored in /root/reference; Pillow 12.2 in this
image): precompute_coeffs / normalize_coeffs_8bpc / ImagingResampleHorizontal_8bpc / ImagingResampleVertical_8bpc of
src/libImaging/Resample.c restated from the published source. Pinned: tests/test_preproc.py checks it bit for bit against
tests/golden/preproc.pt, which oracle/gen_preproc_golden.py produced with the real torchvision + Pillow calls."""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def coeffs(in_size, out_size):
    scale = filterscale = in_size / out_size
    filterscale = max(filterscale, 1.0)
    support = filterscale  # bilinear: support 1.0
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int64)
    kk = np.zeros((out_size, ksize), np.int64)
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = [max(0.0, 1.0 - abs((x + xmin - center + 0.5) / filterscale)) for x in range(xmax)]
        # Pillow multiplies by ss = 1/filterscale instead of dividing; both are restated and must agree on the fixtures
        ss = 1.0 / filterscale
        w = [(1.0 - abs((x + xmin - center + 0.5) * ss)) if abs((x + xmin - center + 0.5) * ss) < 1.0 else 0.0 for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        w = [v / ww if ww != 0.0 else v for v in w]
        bounds[xx] = (xmin, xmax)
        for x, v in enumerate(w):
            kk[xx, x] = int(0.5 + v * (1 << PRECISION_BITS)) if v >= 0 else int(-0.5 + v * (1 << PRECISION_BITS))
    return bounds, kk


def _pass(img, bounds, kk, axis):
    """img uint8 (H, W, 3); resample along `axis` (1 = horizontal, 0 = vertical)."""
    img = np.moveaxis(img.astype(np.int64), axis, 0)
    out = np.empty((bounds.shape[0],) + img.shape[1:], np.int64)
    for o in range(bounds.shape[0]):
        lo, cnt = bounds[o]
        acc = np.full(img.shape[1:], 1 << (PRECISION_BITS - 1), np.int64)
        for x in range(cnt):
            acc += img[lo + x] * kk[o, x]
        out[o] = np.clip(acc >> PRECISION_BITS, 0, 255)
    return np.moveaxis(out, 0, axis).astype(np.uint8)


def transform_u8(frame, crop, resize):
    """frame uint8 (H, W, 3) -> uint8 (OH, OW, 3) as PIL holds it after Crop + Resize."""
    img = frame[crop[0]:frame.shape[0] - crop[1]]
    oh, ow = resize
    hb, hk = coeffs(img.shape[1], ow)
    vb, vk = coeffs(img.shape[0], oh)
    return _pass(_pass(img, hb, hk, 1), vb, vk, 0)


def transform(frame, crop, resize):
    """-> float32 (3, OH, OW) in [0, 1] (ToTensor)."""
    u8 = transform_u8(frame, crop, resize)
    return (u8.transpose(2, 0, 1).astype(np.float32) / np.float32(255))
