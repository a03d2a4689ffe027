"""Input pipeline. CPU: the oracle port of Crop + Pillow bilinear Resize + ToTensor reproduces the
fixtures made with the real torchvision/Pillow calls bit for bit, and the product's coefficient tables equal the oracle's.
GPU: `pmoe_preprocess_frames` through the host wrapper reproduces the same fixtures bit for bit (uint8 stage and float
output), for CUDA and pinned-host inputs and for the (B, T, H, W, 3) -> (B, T, 3, H, W) stacking."""
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "preproc.pt")


def _cases():
    return torch.load(GOLD)["cases"]


def test_oracle_port_matches_torchvision_pillow_fixtures():
    from oracle import preproc_port as P
    for c in _cases():
        for f, want in zip(c["frames"].numpy(), c["out_u8"]):
            got = P.transform_u8(f, c["crop"], c["resize"])
            assert np.array_equal(got.transpose(2, 0, 1), want.numpy())
            assert np.array_equal(P.transform(f, c["crop"], c["resize"]), want.float().div(255).numpy())


def test_product_coefficient_tables_equal_the_oracle():
    from oracle import preproc_port as P
    from pmoe_b200.preproc import pillow_bilinear_coeffs
    for (i, o) in [(800, 224), (385, 224), (224, 224), (100, 224), (333, 80), (7, 5), (5, 7)]:
        b, k, ks = pillow_bilinear_coeffs(i, o)
        ob, ok_ = P.coeffs(i, o)
        assert ks == ok_.shape[1]
        assert np.array_equal(b.numpy(), ob) and np.array_equal(k.numpy(), ok_)
        assert (k.sum(1) - (1 << 22)).abs().max() <= ks  # rows sum to one in fixed point, up to per-tap rounding


@pytest.mark.gpu
def test_gpu_preprocess_is_bit_exact():
    from pmoe_b200.preproc import FramePreprocessor
    for c in _cases():
        pp = FramePreprocessor(c["crop"], c["resize"])
        want_u8 = c["out_u8"]
        want = want_u8.float().div(255)
        got, u8 = pp(c["frames"].cuda(), want_u8=True)
        assert torch.equal(u8.cpu().permute(0, 3, 1, 2), want_u8)
        assert torch.equal(got.cpu(), want)
        got2 = pp(c["frames"].pin_memory())              # pinned host frames: uploaded by the call
        assert torch.equal(got2.cpu(), want)


@pytest.mark.gpu
def test_gpu_preprocess_stacks_frames_like_the_dataset():
    from pmoe_b200.preproc import FramePreprocessor
    c = _cases()[0]
    pp = FramePreprocessor(c["crop"], c["resize"])
    frames = c["frames"]
    seq = torch.stack([frames, frames.flip(0)], 0)          # (B=2, T=2, H, W, 3)
    out = pp(seq.cuda())
    want = c["out_u8"].float().div(255)
    assert out.shape == (2, 2, 3) + tuple(c["resize"])
    assert torch.equal(out[0].cpu(), want) and torch.equal(out[1].cpu(), want.flip(0))
