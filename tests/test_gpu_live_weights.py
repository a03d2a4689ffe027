"""The packed operands and folded statistics the kernels read are DERIVED from the live parameters / buffers: a fused
optimizer step, a BatchNorm statistics update or an SWA update must be visible to the next forward — eagerly and inside a
captured CUDA graph (reference behaviour: ATen reads `conv.weight` / `bn.running_mean` directly on every call,
model/blocks/basics.py:51-55; trainer/train_2.py:157-165,179-187)."""
import copy

import pytest
import torch

from conftest import rel_err
from oracle import functional as O

pytestmark = pytest.mark.gpu
dev = "cuda"


def test_pack_gather_and_scatter_kernels_vs_torch():
    from pmoe_b200 import packs
    g = torch.Generator().manual_seed(3)
    for (cout, cin, k, rows, cols) in ((23, 7, 3, 32, 9 * 16), (64, 64, 3, 64, 9 * 64), (5, 1536, 1, 16, 1536)):
        w = torch.randn(cout, cin, k, k, generator=g).to(dev)
        # a packing recipe with padding on both axes and a permutation
        def build(wf):
            m = wf.permute(0, 2, 3, 1).reshape(cout, -1)                      # (cout, tap*cin)
            m = torch.nn.functional.pad(m, (0, cols - m.shape[1] if cols > m.shape[1] else 0))[:, :cols]
            return torch.nn.functional.pad(m, (0, 0, 0, rows - cout))
        want = build(w)
        for dt in (torch.bfloat16, torch.float32):
            p = torch.nn.Parameter(w.clone())
            got, ent = packs.packed(p, "t%s" % dt, p.shape, build, dt)
            assert torch.equal(got.float(), want.to(dt).float())
            # stale after a raw-pointer style update + version bump
            with torch.no_grad():
                p.data.mul_(2.0)
            packs.bump([p])
            got2, _ = packs.packed(p, "t%s" % dt, p.shape, build, dt)
            assert got2.data_ptr() == got.data_ptr()                              # static buffer
            assert torch.equal(got2.float(), (2.0 * want).to(dt).float())
        # scatter: the packed gradient goes back to the parameter layout; padding is ignored; accumulate adds
        gp = torch.randn(rows, cols, generator=g).to(dev)
        dst = torch.full((w.numel(),), 7.0, device=dev)
        packs.scatter_grad(gp, ent.idx, dst, False)
        idx = ent.idx.reshape(-1).long()
        ref = torch.full((w.numel(),), 7.0, device=dev)
        ref[idx[idx >= 0]] = gp.reshape(-1)[idx >= 0]
        assert torch.equal(dst, ref)
        packs.scatter_grad(gp, ent.idx, dst, True, alpha=0.5)
        ref[idx[idx >= 0]] += 0.5 * gp.reshape(-1)[idx >= 0]
        assert torch.allclose(dst, ref)
        # the inverse-map form the training path uses (coalesced writes), single and grouped, fresh and accumulating
        if cols >= cin * k * k:      # the recipe above truncates columns otherwise: not a bijection, no inverse
            d1 = torch.full((w.numel(),), 3.0, device=dev)
            packs.unpack_grads(gp, ent, [d1], [False])
            want1 = torch.zeros(w.numel(), device=dev)
            want1[idx[idx >= 0]] = gp.reshape(-1)[idx >= 0]
            assert torch.equal(d1, want1)
            gp3 = torch.stack([gp, 2 * gp, 3 * gp]).contiguous()
            ds = [torch.ones(w.numel(), device=dev), None, torch.zeros(w.numel(), device=dev)]
            packs.unpack_grads(gp3, ent, ds, [True, False, False])
            assert torch.equal(ds[0], want1 + 1) and torch.equal(ds[2], 3 * want1)


def _unet_small():
    sd = O.seeded_state_dict(O.make_spec(O.unet_spec, 3, 23), 5)
    gen = torch.Generator().manual_seed(6)
    x = torch.rand(2, 3, 32, 32, generator=gen)
    up = torch.randn(2, 23, 32, 32, generator=gen) * 1e-2
    return sd, x, up


def test_forward_sees_the_fused_adam_step_fp32_vs_torch_adam():
    """forward -> backward -> FusedAdam.step -> forward, against the oracle stepped by torch.optim.Adam: the second
    forward must use the UPDATED conv / convT / 1x1 weights (it used the first step's packed copies before)."""
    from pmoe_b200 import config, optim
    from pmoe_b200.model.blocks.unet import UNet
    sd, x, up = _unet_small()
    leaf = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone()) for k, v in sd.items()}
    params = [v for v in leaf.values() if v.requires_grad]
    ref_opt = torch.optim.Adam(params, lr=1e-3, amsgrad=True)
    outs_ref = []
    for _ in range(3):
        ref_opt.zero_grad()
        o = O.unet(x, leaf, "", True)
        outs_ref.append(o.detach().clone())
        o.backward(up)
        ref_opt.step()
    with config.use_precision("fp32"):
        net = UNet(3, 23)
        net.load_state_dict(sd, strict=True)
        net = net.to(dev).train()
        opt = optim.FusedAdam(net.parameters(), lr=1e-3, amsgrad=True)
        outs = []
        for _ in range(3):
            opt.zero_grad(set_to_none=True)
            o = net(x.to(dev))
            outs.append(o.detach().cpu())
            o.backward(up.to(dev))
            opt.step()
    e = [rel_err(a, b) for a, b in zip(outs, outs_ref)]
    moved = rel_err(outs[1], outs[0])
    print("\n[live weights] step-by-step logits vs oracle+torch.optim.Adam: %s ; step1 vs step0 moved by %.3e" % (["%.2e" % v for v in e], moved))
    assert moved > 1e-2                      # the first Adam step moves every weight by lr: the logits move visibly
    # a stale packed copy would leave e[1] at the size of `moved`; Adam's 1/sqrt(v) amplifies 1e-6 gradient differences on tiny gradients
    assert e[0] < 1e-4 and e[1] < 0.05 * moved and e[2] < 0.5 * moved


def test_cuda_graph_replay_trains_bf16():
    """The captured micro-step re-packs the operands from the live parameters on every replay: replay + FusedAdam must
    follow the same trajectory as the eager tape + FusedAdam, and the loss must fall."""
    from pmoe_b200 import conf, loss as L, optim
    from pmoe_b200.model.moe import get_model
    K, B = 2, 8
    gen = torch.Generator().manual_seed(12)
    d = {"images": torch.rand(B, 4, 3, 64, 64, generator=gen), "speed": torch.rand(B, 1, generator=gen) * 1.2,
         "command": torch.nn.functional.one_hot(torch.randint(0, 6, (B,), generator=gen), 6).float(),
         "control": torch.rand(B, 2, generator=gen) * 2 - 1, "target": torch.rand(B, 1, generator=gen)}
    d = {k: v.to(dev) for k, v in d.items()}
    cfg = conf.stage2_model_cfg("moe", K, dropout=0.0)
    torch.manual_seed(1)
    model0 = get_model(cfg)
    torch.distributions.Distribution.set_default_validate_args(False)

    def trajectory(use_graph, steps=10):
        model = copy.deepcopy(model0).to(dev).train()
        opt = optim.FusedAdam(model.parameters(), lr=1e-4, amsgrad=True)

        def fwd_bwd():
            dist_, sp = model(d["images"], d["speed"], d["command"])
            loss = L.moe_loss(dist_, sp, d["control"], d["target"].clone(), cfg.loss_coefs)
            loss.backward()
            return loss
        losses = []
        if not use_graph:
            for _ in range(steps):
                opt.zero_grad(set_to_none=True)
                losses.append(fwd_bwd().item())
                opt.step(max_grad_norm=1.0)
            return losses, model
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                opt.zero_grad(set_to_none=True)
                fwd_bwd()
        torch.cuda.current_stream().wait_stream(side)
        for p in model.parameters():
            p.grad.zero_()
        with torch.no_grad():  # the warm-up passes updated the BatchNorm statistics: restore them for the comparison
            model.load_state_dict(copy.deepcopy(model0).state_dict())
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            static_loss = fwd_bwd()
        for p in model.parameters():
            p.grad.zero_()
        with torch.no_grad():
            model.load_state_dict(copy.deepcopy(model0).state_dict())
        for _ in range(steps):
            torch._foreach_zero_([p.grad for p in model.parameters()])
            graph.replay()
            losses.append(static_loss.item())
            opt.step(max_grad_norm=1.0)
        return losses, model

    le, me = trajectory(False)
    lg, mg = trajectory(True)
    print("\n[graph training] eager losses %s\n                 graph losses %s" % (["%.4f" % v for v in le], ["%.4f" % v for v in lg]))
    assert min(le[1:]) < le[0] - 1e-3 and min(lg[1:]) < lg[0] - 1e-3  # the network learns (it did not while the packed copies were stale)
    assert len(set("%.5f" % v for v in lg)) > 5                      # every replay sees new weights
    # same trajectory while rounding noise has not been amplified yet (B = 8 at random init is chaotic after a few steps)
    assert max(abs(a - b) / max(1.0, abs(a)) for a, b in zip(le[:4], lg[:4])) < 1e-2
    moved = max(rel_err(p.detach().cpu(), q.detach()) for p, q in zip(mg.parameters(), model0.parameters()) if q.numel() > 1000)
    assert moved > 1e-4
    torch.distributions.Distribution.set_default_validate_args(True)


def test_eval_follows_bn_statistics_updates():
    """train -> eval -> train -> eval: the eval passes fold the CURRENT running statistics (the fold cache is keyed on the
    buffers' versions, which the statistics kernel now bumps)."""
    from pmoe_b200 import config
    from pmoe_b200.model.blocks.unet import UNet
    sd, x, _ = _unet_small()
    x2 = torch.rand(2, 3, 32, 32, generator=torch.Generator().manual_seed(60)) * 3.0
    ref_sd = {k: v.clone() for k, v in sd.items()}
    ref_evals = []
    with torch.no_grad():
        for xin in (x, x2):
            O.unet(xin, ref_sd, "", True)                 # updates the running statistics in ref_sd
            ref_evals.append(O.unet(x, ref_sd, "", False))
    with config.use_precision("fp32"):
        net = UNet(3, 23)
        net.load_state_dict(sd, strict=True)
        net = net.to(dev)
        evals = []
        with torch.no_grad():
            for xin in (x, x2):
                net.train()
                net(xin.to(dev))
                net.eval()
                evals.append(net(x.to(dev)).cpu())
    e = [rel_err(a, b) for a, b in zip(evals, ref_evals)]
    changed = rel_err(evals[1], evals[0])
    print("\n[bn fold] eval after 1 and 2 statistics updates vs oracle: %.2e %.2e ; eval moved by %.2e" % (e[0], e[1], changed))
    assert changed > 1e-3
    assert max(e) < 1e-4
    assert int(net.state_dict()["dwn_1.1.num_batches_tracked"]) == 2


def test_swa_model_buffers_and_parameters_vs_torch():
    from pmoe_b200 import config, optim
    from pmoe_b200.model.blocks.unet import UNet
    sd, x, _ = _unet_small()
    with config.use_precision("fp32"):
        net = UNet(3, 23)
        net.load_state_dict(sd, strict=True)
        net = net.to(dev).train()
        ours = optim.AveragedModel(net)
        theirs = torch.optim.swa_utils.AveragedModel(net)
        for i in range(3):
            with torch.no_grad():
                net(x.to(dev) * (i + 1))                   # moves the BatchNorm statistics
                for p in net.parameters():
                    p.add_(0.01 * (i + 1))
            from pmoe_b200 import packs
            packs.bump(list(net.parameters()))
            ours.update_parameters(net)
            theirs.update_parameters(net)
    a, b = ours.state_dict(), theirs.state_dict()
    assert a.keys() == b.keys()
    for k in a:
        assert torch.allclose(a[k].float(), b[k].float(), rtol=1e-6, atol=1e-7), k
    assert int(a["module.dwn_1.1.num_batches_tracked"]) == 3


def test_fused_adam_per_parameter_step_counters_vs_torch():
    """A parameter that starts receiving gradients later has its own bias correction (torch.optim.Adam semantics)."""
    from pmoe_b200 import optim
    g = torch.Generator().manual_seed(2)
    ws = [torch.randn(300, generator=g), torch.randn(70, 9, generator=g)]
    pa = [torch.nn.Parameter(w.clone().to(dev)) for w in ws]
    pb = [torch.nn.Parameter(w.clone().to(dev)) for w in ws]
    oa, ob = optim.FusedAdam(pa, lr=1e-2, amsgrad=True), torch.optim.Adam(pb, lr=1e-2, amsgrad=True)
    for step in range(5):
        for plist in (pa, pb):
            gg = torch.Generator().manual_seed(100 + step)
            for i, p in enumerate(plist):
                p.grad = None if (i == 1 and step < 2) else torch.randn(p.shape, generator=gg).to(dev)
        oa.step()
        ob.step()
    for p, q in zip(pa, pb):
        assert torch.allclose(p, q, rtol=1e-5, atol=1e-7)
    assert pa[0]._version >= 5           # the fused step bumps the version counters


def test_sample_consumes_the_generator_like_the_reference():
    """`sample()` draws with the same torch.distributions objects as the reference (moe.py:160-177): under the same CUDA
    generator state the component index and the sample are bit-identical to MixtureSameFamily.sample() on the model's own
    (probs, mean, std)."""
    import torch.distributions as D
    from pmoe_b200 import conf
    from pmoe_b200.model.moe import get_model
    torch.manual_seed(3)
    cfg = conf.stage2_model_cfg("moe", 3, dropout=0.0)
    model = get_model(cfg).to(dev).eval()
    gen = torch.Generator().manual_seed(4)
    B = 16
    images, speed = torch.rand(B, 4, 3, 64, 64, generator=gen).to(dev), (torch.rand(B, 1, generator=gen) * 1.2).to(dev)
    command = torch.nn.functional.one_hot(torch.randint(0, 6, (B,), generator=gen), 6).float().to(dev)
    with torch.no_grad():
        probs, mean, std, _, route = model.components(images, speed, command)
        torch.manual_seed(99)
        got = model.sample(images, speed, command)
        torch.manual_seed(99)
        mix, comp = D.Categorical(probs), D.Independent(D.Normal(mean, std), 1)
        want = D.MixtureSameFamily(mix, comp).sample()
        torch.manual_seed(99)
        k = D.Categorical(probs).sample()               # the component index the draw used
    assert torch.equal(got, want)
    picked = mean.gather(1, k.view(B, 1, 1).expand(B, 1, 2)).squeeze(1)
    assert ((got - picked).abs() <= 6.0 * std.gather(1, k.view(B, 1, 1).expand(B, 1, 2)).squeeze(1)).all()
    assert torch.equal(route, probs.argmax(1))
