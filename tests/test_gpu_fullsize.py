"""Parity at BASELINE.json's full sizes through size-independent properties (the oracle cannot run these sizes in seconds):

* linearity of the tensor-core convolution and of its weight gradient under a power-of-two scaling of the input — exact in
  bf16/fp32, so the comparison is bit for bit — at 224^2 / 112^2 with the bench's channel counts (resident and streamed
  halo kernels, TC wgrad);
* batch independence and batch-permutation equivariance of the PU-Net forward at configs[1]'s B = 256: an image's logits do
  not depend on its position in the batch or on its neighbours (eval mode). Tile boundaries and the order of the per-image
  ECA atomics move with the position, so the comparison uses the bf16 tolerance of north_star (1e-2 relative) and
  the predicted class map;
* the small-batch run of the same images agrees with the CPU oracle (fp32) within the bf16 tolerance, which anchors the
  full-size run to the reference by transitivity.
"""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pytestmark = pytest.mark.gpu
dev = "cuda"


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("B,H,cin,cout", [(64, 224, 64, 64), (64, 112, 256, 128), (64, 112, 128, 128)])
def test_conv_linearity_is_exact_at_bench_sizes(B, H, cin, cout):
    from pmoe_b200 import ops
    g = torch.Generator().manual_seed(3)
    x = torch.randn(B, H, H, cin, generator=g).to(dev).to(torch.bfloat16)
    w = (torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5).to(dev)
    cp, cop = [cin], ops.cout_padded(cout)
    ck = ops.choose_ck(cp)
    wp = ops.pack_conv_weight(w, [cin], cp, ops.TAPS3, cop)
    segs = ops.conv_segments([(r - 1, s - 1) for (r, s) in ops.TAPS3], cp, ck)
    y1 = torch.empty(B, H, H, ops.pad_ch(cout), dtype=torch.bfloat16, device=dev)
    y2 = torch.empty_like(y1)
    ops.conv_tc([x], wp, segs, ck, y1, None, None, "none")
    ops.conv_tc([x * 4], wp, segs, ck, y2, None, None, "none")
    assert torch.isfinite(y1.float()).all() and y1.float().abs().max() > 0.1
    assert torch.equal(y2, y1 * 4)
    # spot check against torch on one image (fp32 math on the same bf16 operands)
    ref = torch.nn.functional.conv2d(x[:1].float().permute(0, 3, 1, 2), w.to(torch.bfloat16).float(), padding=1)
    assert _rel(y1[:1, ..., :cout].float().permute(0, 3, 1, 2), ref) < 1e-2
    # weight gradient: linear in dy, exact under a power-of-two scaling; fp32 atomics reorder sums, hence 1e-5 not equality
    dy = torch.randn(B, H, H, ops.pad_ch(cout), generator=g).to(dev).to(torch.bfloat16)
    dw1 = torch.zeros(cop, wp.shape[1], dtype=torch.float32, device=dev)
    dw2 = torch.zeros_like(dw1)
    ops.conv_wgrad([x], segs, ck, dy, dw1)
    ops.conv_wgrad([x], segs, ck, dy * 2, dw2)
    assert _rel(dw2, dw1 * 2) < 1e-5


def test_punet_b256_batch_independence_and_oracle_anchor():
    import bench
    from oracle import functional as O
    torch.manual_seed(0)
    net = bench.build_punet().to(dev).eval()
    B = 256
    imgs = bench.synth_images(B, seed=77)
    with torch.no_grad():
        full = net(imgs.to(dev))                                   # configs[1]: (256, 6, 23, 224, 224) fp32
        assert full.shape == (B, 6, 23, 224, 224) and torch.isfinite(full).all()
        pick = torch.tensor([0, 5, 131, 255])
        small = net(imgs[pick].to(dev))                            # the same four images on their own
        perm = torch.randperm(B, generator=torch.Generator().manual_seed(1))
        permuted = net(imgs[perm].to(dev))
    e_small = _rel(full[pick.to(dev)], small)
    e_perm = _rel(permuted, full[perm.to(dev)])
    agree = (full[pick.to(dev)].argmax(2) == small.argmax(2)).float().mean().item()
    print("\n[B=256] batch independence rel %.3e, permutation rel %.3e, class-map agreement %.5f" % (e_small, e_perm, agree))
    assert e_small < 1e-2 and e_perm < 1e-2   # north_star bf16 tolerance
    assert agree > 0.995
    # anchor: one of the picked images through the fp32 CPU oracle (seconds on the host cores)
    sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
    with torch.no_grad():
        ref = O.punet(imgs[pick[:1]], sd, "", False, 4, 6)
    e_ref = _rel(small[:1].cpu(), ref)
    print("[B=256] oracle anchor rel %.3e" % e_ref)
    assert e_ref < 1e-2


def test_moe_train_step_properties_at_config2_micro_batch():
    """BASELINE configs[2]: 6-expert mixture, one micro-batch of 128 samples, bf16 tensor-core path. Properties that do not
    need an oracle at this size: the gating weights are a distribution over the experts, the routing index is in range and
    equals the arg-max of the weights, every parameter receives a finite gradient, and the step is invariant to the order of
    the samples in the batch (BatchNorm statistics and the mean loss are permutation invariant; summation order moves,
    so the comparison is at the bf16 tolerance)."""
    from pmoe_b200 import conf, loss as L
    from pmoe_b200.model.moe import get_model
    K, B = 6, 128
    torch.manual_seed(0)
    cfg = conf.stage2_model_cfg("moe", K, dropout=0.0)   # no dropout: the two runs must see the same network
    model = get_model(cfg).to(dev).train()
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    g = torch.Generator().manual_seed(9)
    d = {"images": torch.rand(B, 4, 3, 224, 224, generator=g), "speed": torch.rand(B, 1, generator=g) * 1.2,
         "command": torch.nn.functional.one_hot(torch.randint(0, 6, (B,), generator=g), 6).float(),
         "control": torch.rand(B, 2, generator=g) * 2 - 1, "target": torch.rand(B, 1, generator=g)}

    def run(order):
        model.load_state_dict(sd0)
        for p in model.parameters():
            p.grad = None
        x = {k: v[order].to(dev) for k, v in d.items()}
        dist_, sp = model(x["images"], x["speed"], x["command"])
        loss = L.moe_loss(dist_, sp, x["control"], x["target"], cfg.loss_coefs)
        loss.backward()
        return dist_, sp, loss.detach(), {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}

    ident = torch.arange(B)
    dist_, sp, loss, grads = run(ident)
    w = dist_.mixture_distribution.probs
    assert w.shape == (B, K) and torch.isfinite(w).all()
    assert (w.sum(1) - 1).abs().max().item() < 1e-5 and (w >= 0).all()
    assert torch.isfinite(loss) and len(grads) > 100
    assert all(torch.isfinite(v).all() for v in grads.values())
    assert sum(int(v.abs().sum() > 0) for v in grads.values()) > 0.9 * len(grads)
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(2))
    dist2, sp2, loss2, grads2 = run(perm)
    w2 = dist2.mixture_distribution.probs
    assert _rel(w2, w[perm.to(dev)]) < 1e-2
    assert abs(loss2.item() - loss.item()) < 1e-2 * max(1.0, abs(loss.item()))
    assert (w2.argmax(1) == w[perm.to(dev)].argmax(1)).float().mean().item() > 0.97   # routing follows the samples
