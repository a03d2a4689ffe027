"""GPU parity of the optimizer-side multi-tensor kernels against what the reference's loop calls on the same
values: torch.nn.utils.clip_grad_norm_, utils.nn.check_grad_norm (train_2.py:160-163) and torch.optim.Adam(amsgrad)
(conf/stage_2.yaml:137-144). fp32 elementwise arithmetic: tolerance 1e-6 relative."""
import copy

import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu
SHAPES = [(64, 12, 3, 3), (64,), (512, 1536), (1, 1, 3), (4, 512), (70001,), (3,)]


def make_params(seed, dev):
    g = torch.Generator().manual_seed(seed)
    ps = [torch.nn.Parameter(torch.randn(*s, generator=g).to(dev)) for s in SHAPES]
    for p in ps:
        p.grad = (torch.randn(*p.shape, generator=g) * 3.0).to(dev)
    return ps


def test_clip_and_check_grad_norm_match_torch():
    from pmoe_b200 import optim
    a, b = make_params(1, "cuda"), make_params(1, "cpu")
    want64 = sum((p.grad.double() ** 2).sum() for p in b).sqrt().item()  # exact norm; torch's fp32 norm-of-norms is ~1e-5 off it
    want = torch.nn.utils.clip_grad_norm_(b, 1.0)
    lin = torch.nn.Module()
    lin.ps = torch.nn.ParameterList(a)
    before = optim.check_grad_norm(lin)
    got = optim.clip_grad_norm_(a, 1.0)
    assert abs(before - want64) / want64 < 1e-6
    assert abs(before - want.item()) / want.item() < 2e-5
    assert abs(got.item() - want64) / want64 < 1e-6
    for pa, pb in zip(a, b):
        assert rel_err(pa.grad.cpu(), pb.grad) < 2e-5
    assert abs(optim.check_grad_norm(lin) - 1.0) < 1e-5
    # a norm below the threshold leaves gradients untouched
    g0 = [p.grad.clone() for p in a]
    optim.clip_grad_norm_(a, 10.0)
    assert all(torch.equal(x, p.grad) for x, p in zip(g0, a))


@pytest.mark.parametrize("amsgrad,wd", [(True, 0.0), (False, 0.0), (True, 1e-2)])
def test_fused_adam_matches_torch_adam(amsgrad, wd):
    from pmoe_b200 import optim
    a, b = make_params(2, "cuda"), make_params(2, "cpu")
    oa = optim.FusedAdam(a, lr=2e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=wd, amsgrad=amsgrad)
    ob = torch.optim.Adam(b, lr=2e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=wd, amsgrad=amsgrad)
    g = torch.Generator().manual_seed(9)
    for step in range(4):
        for pa, pb in zip(a, b):
            gr = torch.randn(*pb.shape, generator=g) * (0.5 + step)
            pb.grad = gr.clone()
            pa.grad = gr.cuda()
        oa.step()
        ob.step()
    for pa, pb in zip(a, b):
        assert rel_err(pa.detach().cpu(), pb.detach()) < 1e-6
        sa, sb = oa.state[pa], ob.state[pb]
        assert rel_err(sa["exp_avg"].cpu(), sb["exp_avg"]) < 1e-6
        assert rel_err(sa["exp_avg_sq"].cpu(), sb["exp_avg_sq"]) < 1e-6
        if amsgrad:
            assert rel_err(sa["max_exp_avg_sq"].cpu(), sb["max_exp_avg_sq"]) < 1e-6
    # state_dict layout interchanges with torch.optim.Adam
    sd = oa.state_dict()
    ob2 = torch.optim.Adam(make_params(2, "cuda"), lr=2e-4, amsgrad=amsgrad)
    ob2.load_state_dict(copy.deepcopy(sd))
    assert set(sd["state"][0].keys()) == ({"step", "exp_avg", "exp_avg_sq"} | ({"max_exp_avg_sq"} if amsgrad else set()))


def test_fused_adam_with_folded_clip_equals_clip_then_step():
    from pmoe_b200 import optim
    a, b = make_params(3, "cuda"), make_params(3, "cpu")
    oa = optim.FusedAdam(a, lr=1e-3, amsgrad=True)
    ob = torch.optim.Adam(b, lr=1e-3, amsgrad=True)
    torch.nn.utils.clip_grad_norm_(b, 1.0)
    ob.step()
    oa.step(max_grad_norm=1.0)
    for pa, pb in zip(a, b):
        assert rel_err(pa.detach().cpu(), pb.detach()) < 1e-6


@pytest.mark.parametrize("centered,momentum,wd", [(True, 0.0, 0.0), (False, 0.9, 1e-3), (True, 0.5, 1e-4)])
def test_fused_rmsprop_matches_torch_rmsprop(centered, momentum, wd):
    """conf/stage_2.yaml:147-153 (centered, alpha 0.99, momentum 0) and the other branches of torch.optim.RMSprop."""
    from pmoe_b200.optim import FusedRMSprop
    g = torch.Generator().manual_seed(3)
    shapes = [(257,), (64, 33), (3, 5, 7), (100000,)]
    ps_a = [torch.nn.Parameter(torch.randn(s, generator=g).cuda()) for s in shapes]
    ps_b = [torch.nn.Parameter(p.detach().clone()) for p in ps_a]
    kw = dict(lr=2e-4, alpha=0.99, eps=1e-8, weight_decay=wd, momentum=momentum, centered=centered)
    oa, ob = FusedRMSprop(ps_a, **kw), torch.optim.RMSprop(ps_b, **kw)
    for it in range(5):
        for a, b in zip(ps_a, ps_b):
            gr = torch.randn(a.shape, generator=g).cuda() * (0.1 + it)
            a.grad, b.grad = gr.clone(), gr.clone()
        oa.step()
        ob.step()
    for a, b in zip(ps_a, ps_b):
        assert ((a - b).abs().max() / b.abs().max()).item() < 1e-6
    for a, b in zip(ps_a, ps_b):
        for k in ("square_avg", "grad_avg", "momentum_buffer"):
            if k in ob.state[b]:
                ref = ob.state[b][k]
                assert ((oa.state[a][k] - ref).abs().max() / ref.abs().max().clamp_min(1e-12)).item() < 1e-5, k


def test_fused_averaged_model_matches_torch_swa():
    """train_2.py:119-121,179-187: AveragedModel(model) + update_parameters once per epoch."""
    from pmoe_b200.optim import AveragedModel
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(37, 53), torch.nn.ReLU(), torch.nn.Linear(53, 11)).cuda()
    mine, ref = AveragedModel(net), torch.optim.swa_utils.AveragedModel(net)
    for it in range(4):
        with torch.no_grad():
            for p in net.parameters():
                p.add_(torch.randn_like(p) * 0.1)
        mine.update_parameters(net)
        ref.update_parameters(net)
    assert int(mine.n_averaged) == int(ref.n_averaged) == 4
    for a, b in zip(mine.module.parameters(), ref.module.parameters()):
        assert ((a - b).abs().max() / b.abs().max()).item() < 1e-6
