"""B = 1 deployment path: the captured-graph tick equals the eager computation on the same frames —
window bookkeeping (first frame fills the window, then a sliding window of the last T frames), Pillow-exact preprocessing,
policy call — and the eager tick equals building the tensors the way ImageAgent.run_step does (image_agent.py:127-160)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _model():
    from pmoe_b200 import conf
    from pmoe_b200.model.moe import get_model
    torch.manual_seed(0)
    return get_model(conf.stage2_model_cfg("moe", 3, dropout=0.0)).cuda().eval()


def _policy(model):
    # deterministic stand-in for model.sample (which draws from the mixture): mean of the most probable expert + weights
    def f(images, speed, command):
        probs, mean, std, _, _ = model.components(images, speed, command)
        return torch.cat([mean.reshape(1, -1), probs.reshape(1, -1)], 1)
    return f


def test_graph_tick_matches_eager_and_the_agent_recipe():
    from pmoe_b200.agent import RealtimeSampler
    from pmoe_b200.preproc import FramePreprocessor
    model = _model()
    pol = _policy(model)
    hs, ws = 300, 400
    g = torch.Generator().manual_seed(4)
    frames = [torch.randint(0, 256, (hs, ws, 3), generator=g, dtype=torch.uint8) for _ in range(7)]
    speeds = [0.1 * i for i in range(7)]
    cmds = [i % 6 for i in range(7)]
    a = RealtimeSampler(model, (hs, ws), policy=pol, graph=True)
    b = RealtimeSampler(model, (hs, ws), policy=pol, graph=False)
    pp = FramePreprocessor((125, 90), (224, 224))
    window = []
    for i, (f, s, c) in enumerate(zip(frames, speeds, cmds)):
        ga, ea = a.step(f.numpy(), s, c), b.step(f.numpy(), s, c)
        # the agent's recipe: transform the frame, deque(maxlen=T), stack, unsqueeze (image_agent.py:131-156)
        t = pp(f.unsqueeze(0).cuda())[0]
        window = ([t] * 4) if i == 0 else (window[1:] + [t])
        images = torch.stack(window, 0).unsqueeze(0)
        speed = torch.tensor([[s]], dtype=torch.float32, device="cuda")
        command = torch.zeros(1, 6, device="cuda")
        command[0, c] = 1
        with torch.no_grad():
            ref = pol(images, speed, command).reshape(-1).cpu()
        assert torch.allclose(ea, ref, rtol=1e-3, atol=1e-4), (i, (ea - ref).abs().max())
        assert torch.allclose(ga, ea, rtol=1e-3, atol=1e-4), (i, (ga - ea).abs().max())


def test_sample_runs_under_the_graph():
    from pmoe_b200.agent import RealtimeSampler
    model = _model()
    s = RealtimeSampler(model, (240, 320), graph=True)
    g = torch.Generator().manual_seed(5)
    acts = [s.step(torch.randint(0, 256, (240, 320, 3), generator=g, dtype=torch.uint8).numpy(), 0.3, 2) for _ in range(4)]
    assert all(a.shape == (2,) and torch.isfinite(a).all() for a in acts)
    assert not torch.equal(acts[-1], acts[-2])      # fresh random draws on every replay
