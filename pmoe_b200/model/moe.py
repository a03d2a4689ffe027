"""Stage-2 models with the reference's constructor arguments, state_dict keys and forward/sample signatures
(reference: PMoE/model/moe.py). `cfg` is the `model` sub-tree of conf/stage_2*.yaml (attribute + mapping access).

Every forward builds ONE tape (pmoe_b200.train) for the whole model: K expert encoders, head MLPs, the gating
kernel — so backward is one replay of that tape.
"""
import torch
import torch.nn as nn
import torch.distributions as D

from .. import nhwc, train
from ..ops import pad_ch
from ..utils.nn import freeze
from .blocks.backbone import get_backbone, get_unet
from .blocks.basics import make_mlp
from .punet import PredictiveUnet


def get_model(cfg):
    model_type = cfg.type
    assert model_type is not None, "Network type can not be None"
    if model_type in ["moe", "moe_alt"]:
        return MixtureOfExperts(cfg)
    elif model_type in ["moe_shared"]:
        return MixtureOfExpertsShared(cfg)
    elif model_type in ["punet", "punet_inter"]:
        return PUNetExpert(cfg)
    elif model_type in ["pmoe", "pmoe+pretrained"]:
        assert cfg.pmoe.moe_dir != "", "MoE pretrained weights directory should be specified"
        if model_type == "pmoe+pretrained":
            assert cfg.pmoe.punet_dir != "", "PU-Net pretrained weights directory should be specified"
        return PMoE(cfg)
    raise ValueError(f"{model_type} is UNKNOWN, model type should be one of 'moe', 'punet', 'punet_inter', 'pmoe', "
                     f"'pmoe+pretrained', 'moe_alt'")


def _backbone_from(params, n_frames=None, n_channels=3):
    if params.backbone.type != "rgb":
        return get_unet(**{**params.backbone.segmentation, "n_frames": params.backbone.n_frames})
    kw = {**params.backbone.rgb, "n_frames": params.backbone.n_frames if n_frames is None else n_frames}
    if n_channels != 3:
        kw["n_channels"] = n_channels
    return get_backbone(**kw)


def _mixture(probs, mean, std):
    return D.MixtureSameFamily(D.Categorical(probs), D.Independent(D.Normal(mean, std), 1))


@torch.no_grad()
def draw(probs, mean, std):
    """`MixtureSameFamily(...).sample()` exactly as the reference calls it (moe.py:160-177, 352): the same torch.distributions
    objects on the same device, so the generator is consumed identically (Categorical.sample(), then Normal.sample()). While a CUDA
    graph is being captured (the B = 1 agent tick) torch.distributions cannot run — its argument checks synchronise — and the
    graph-safe `sample_mixture` draws from the same distribution instead."""
    if probs.is_cuda and torch.cuda.is_current_stream_capturing():
        return sample_mixture(probs, mean, std)
    return _mixture(probs, mean, std).sample()


@torch.no_grad()
def sample_mixture(probs, mean, std):
    """One draw per row from MixtureSameFamily(Categorical(probs), Independent(Normal(mean, std), 1)) — what the reference's
    `sample()` returns (moe.py:127-133) — written without torch.distributions: component k by inverse CDF of one uniform,
    then mean[k] + std[k] * eps. Same distribution, but no host synchronisation (torch.normal validates `std >= 0` with an
    `.item()`, which also forbids CUDA-graph capture of the B = 1 agent tick)."""
    B, K = probs.shape
    u = torch.rand(B, 1, device=probs.device, dtype=probs.dtype)
    k = (u >= probs.cumsum(-1)).sum(-1).clamp_(max=K - 1)                      # (B,)
    idx = k.view(B, 1, 1).expand(B, 1, mean.shape[-1])
    m, s = mean.gather(1, idx).squeeze(1), std.gather(1, idx).squeeze(1)
    return m + s * torch.randn_like(m)


class BaseExpert(nn.Module):
    """Parameter container of one expert (moe.py:50-72). Runs only inside a mixture's tape."""
    alt = False

    def __init__(self, params):
        super().__init__()
        self.speed_encoder = make_mlp(**params.speed_encoder)
        self.command_encoder = make_mlp(**params.command_encoder)
        self.backbone = _backbone_from(params)
        self.speed_pred = make_mlp(**params.speed_prediction)
        self.action_features = make_mlp(**params.action_head)
        d = params.action_head.dims[-1]
        self.alpha = nn.Linear(d, 1)
        self.action_pred = nn.Linear(d, 4)

    def forward(self, images, speed, command):
        probs, mean, std, speeds, _ = _run_experts(self, [self], images, speed, command, self.alt, softmax=False)
        return probs, mean[:, 0], std[:, 0], speeds[:, 0]


class BaseExpertAlt(BaseExpert):
    alt = True

    def __init__(self, params):
        super().__init__(params)
        self.alpha = nn.Sequential(nn.Linear(1536, 512), nn.ReLU(inplace=True), nn.Linear(512, 1))


def _run_experts(owner, experts, images, speed, command, alt, softmax=True):
    K = len(experts)
    B = images.shape[0]

    def runner(tape):
        x = nhwc.from_nchw(images.reshape(B, -1, images.shape[-2], images.shape[-1]), dtype=tape.dtype)
        speed_a = train.vec_act(tape, speed.reshape(B, -1).float(), speed.reshape(B, -1).shape[1])
        cmd_a = train.vec_act(tape, command.reshape(B, -1).float(), command.reshape(B, -1).shape[1])
        dev = images.device
        alpha_buf = torch.zeros(1, 1, B, K * 16, dtype=tape.dtype, device=dev)
        ap_buf = torch.zeros(1, 1, B, K * 16, dtype=tape.dtype, device=dev)
        sp_buf = torch.zeros(1, 1, B, K * 16, dtype=tape.dtype, device=dev)
        if train.grouped_supported(experts, alt):
            # encoders per expert, then ALL experts' heads with one launch per layer (grouped GEMM over the expert axis)
            feats = []
            small = B * images.shape[-2] * images.shape[-1] <= train.MULTI_STREAM_MAX_PIXELS
            for k, ex in enumerate(experts):   # the encoders share only the input: each on its own stream (train.Tape.branch)
                with tape.branch(k, enable=small and K > 1):
                    feats.append(train.backbone_features(tape, ex.backbone, x, tag="moe.%d.backbone" % k))
            tape.join()
            al, ap, sp = train.expert_heads_grouped(tape, experts, feats, speed_a, cmd_a, alt)
            gm = train.GateMixture(tape, [al], [ap], al.t, ap.t, B, K, relu_alpha=not alt, a_sk=B * 16, p_sk=B * 16)
            speeds = sp.t.view(K, B, 16)[:, :, :1].permute(1, 0, 2).float().contiguous()

            def seed(tp, g):
                gm.backward(g[0], g[1], g[2])
                if g[3] is not None:
                    train.seed_stacked(tp, sp, g[3])
            return [gm.probs, gm.mean, gm.std, speeds, gm.route], seed
        als, aps, sps = [], [], []
        small = B * images.shape[-2] * images.shape[-1] <= train.MULTI_STREAM_MAX_PIXELS
        for k, ex in enumerate(experts):
            with tape.branch(k, enable=small and K > 1):   # encoder + heads of one expert: independent of the others up to the gating
                fa = train.backbone_features(tape, ex.backbone, x, tag="moe.%d.backbone" % k)
                sl = slice(k * 16, (k + 1) * 16)
                al, ap, sp = train.expert_heads(tape, ex, fa, speed_a, cmd_a, alt, alpha_buf[..., sl], ap_buf[..., sl], sp_buf[..., sl],
                                                tag="moe.%d" % k)
            als.append(al)
            aps.append(ap)
            sps.append(sp)
        tape.join()
        gm = train.GateMixture(tape, als, aps, alpha_buf, ap_buf, B, K, relu_alpha=not alt, a_sk=16, p_sk=16)
        speeds = sp_buf.view(B, K, 16)[:, :, :1].float()

        def seed(tp, g):
            gm.backward(g[0], g[1], g[2])
            if g[3] is not None:
                for k, sp in enumerate(sps):
                    train.seed_vec(tp, sp, g[3][:, k, :])
        return [gm.probs, gm.mean, gm.std, speeds, gm.route], seed

    return train.run(owner, runner)


class MixtureOfExperts(nn.Module):
    def __init__(self, params):
        super().__init__()
        self.k = params.n_experts
        base = BaseExpert if params.type == "moe" else BaseExpertAlt  # moe.py:136 (PMoE's inner mixture is Alt too)
        self.alt = base is BaseExpertAlt
        self.moe = nn.ModuleList([base(params) for _ in range(self.k)])

    def components(self, images, speed, command):
        """(probs (B,K), mean (B,K,2), std (B,K,2), speeds (B,K,1), routing index (B,))"""
        return _run_experts(self, list(self.moe), images, speed, command, self.alt)

    def forward(self, images, speed, command):
        probs, mean, std, speeds, _ = self.components(images, speed, command)
        return _mixture(probs, mean, std), speeds

    def sample(self, images, speed, command) -> torch.Tensor:
        probs, mean, std, _, _ = self.components(images, speed, command)
        return draw(probs, mean, std)


class MixtureOfExpertsShared(nn.Module):
    """Shared encoder, Linear gating (moe.py:180-265)."""

    def __init__(self, params):
        super().__init__()
        self.speed_encoder = make_mlp(**params.speed_encoder)
        self.command_encoder = make_mlp(**params.command_encoder)
        self.backbone = _backbone_from(params)
        self.speed_pred = make_mlp(**params.speed_prediction)
        self.action_features = make_mlp(**params.action_head)
        d = params.action_head.dims[-1]
        self.n_experts = params.n_experts
        self.alpha = nn.Linear(d, params.n_experts)
        self.action_pred = nn.Linear(d, 4 * params.n_experts)

    def components(self, images, speed, command):
        K, B = self.n_experts, images.shape[0]

        def runner(tape):
            x = nhwc.from_nchw(images.reshape(B, -1, images.shape[-2], images.shape[-1]), dtype=tape.dtype)
            speed_a = train.vec_act(tape, speed.reshape(B, -1).float(), 1)
            cmd_a = train.vec_act(tape, command.reshape(B, -1).float(), command.reshape(B, -1).shape[1])
            feat = train.backbone_features(tape, self.backbone, x)
            s = train.mlp(tape, self.speed_encoder, [speed_a], "speed_encoder")
            c = train.mlp(tape, self.command_encoder, [cmd_a], "command_encoder")
            feats = [feat, s, c]
            sp = train.mlp(tape, self.speed_pred, feats, "speed_pred")
            af = train.mlp(tape, self.action_features, feats, "action_features")
            ap = train.linear_op(tape, [af], self.action_pred, None, tag="action_pred")
            al = train.linear_op(tape, [af], self.alpha, None, tag="alpha")
            gm = train.GateMixture(tape, [al], [ap], al.t, ap.t, B, K, relu_alpha=False, a_sk=1, p_sk=4)

            def seed(tp, g):
                gm.backward(g[0], g[1], g[2])
                train.seed_vec(tp, sp, g[3])
            return [gm.probs, gm.mean, gm.std, train.vec_value(sp), gm.route], seed

        return train.run(self, runner)

    def forward(self, images, speed, command):
        probs, mean, std, speed_pred, _ = self.components(images, speed, command)
        return _mixture(probs, mean, std), speed_pred

    def sample(self, images, speed, command) -> torch.Tensor:
        probs, mean, std, _, _ = self.components(images, speed, command)
        return draw(probs, mean, std)


class PUNetExpert(nn.Module):
    """PU-Net as action predictor (moe.py:268-323), including the constructor's checkpoint loads and freeze."""

    def __init__(self, params):
        super().__init__()
        self.return_inter = True if params.type == "punet_inter" else False
        params.punet.inter_repr = self.return_inter  # the reference mutates the config too (moe.py:274)
        self.speed_encoder = make_mlp(**params.speed_encoder)
        self.command_encoder = make_mlp(**params.command_encoder)
        self.punet = PredictiveUnet(**params.punet)
        punet_weights = torch.load(params.punet_path, map_location=params.device)
        self.punet.load_state_dict(punet_weights["model"])
        self.punet = freeze(self.punet)
        self.backbone = None if self.return_inter else _backbone_from(params, params.punet.future_frames, params.punet.num_classes)
        self.speed_pred = make_mlp(**params.speed_prediction)
        self.action_pred = nn.Sequential(make_mlp(**params.action_head), nn.Linear(params.action_head.dims[-1], 2))

    def forward(self, images, speed, command):
        B = images.shape[0]

        def runner(tape):
            speed_a = train.vec_act(tape, speed.reshape(B, -1).float(), 1)
            cmd_a = train.vec_act(tape, command.reshape(B, -1).float(), command.reshape(B, -1).shape[1])
            s = train.mlp(tape, self.speed_encoder, [speed_a], "speed_encoder")
            c = train.mlp(tape, self.command_encoder, [cmd_a], "command_encoder")
            r = train.punet_tape(tape, self.punet, images)
            if self.return_inter:
                img = train.feature_act(tape, r["inter"])
            else:
                P, Fu, slot, ncls = r["P"], r["F"], r["slot"], r["ncls"]
                fut = train.ring_window(tape, r["ring"], r["futures"], P * slot, slot, ncls)
                if hasattr(self.backbone, "tape_features"):   # MobileNet family
                    img = self.backbone.tape_features(tape, fut, "backbone", (Fu, ncls, slot), r["pools"][:, P * slot:(P + Fu) * slot])
                else:
                    stem = train.eca_conv_block(tape, self.backbone.conv1, fut, (Fu, ncls, slot), r["pools"][:, P * slot:(P + Fu) * slot],
                                                tag="backbone.conv1", want_out_stats=self.backbone.bn1.training)
                    img = train.backbone_head(tape, self.backbone, train.resnet18_after_stem(tape, self.backbone, stem))
            feats = [img, s, c]
            af = train.mlp(tape, self.action_pred[0], feats, "action_pred.0")
            act = train.linear_op(tape, [af], self.action_pred[1], "tanh", tag="action_pred.1")
            sp = train.mlp(tape, self.speed_pred, feats, "speed_pred")

            def seed(tp, g):
                train.seed_vec(tp, act, g[0])
                train.seed_vec(tp, sp, g[1])
            return [train.vec_value(act), train.vec_value(sp)], seed

        out = train.run(self, runner)
        return out[0], out[1]

    def sample(self, images, speed, command) -> torch.Tensor:
        action, _ = self.forward(images, speed, command)
        return action


class PMoE(nn.Module):
    """Predictive mixture of experts (moe.py:326-363)."""

    def __init__(self, params):
        super().__init__()
        assert params.pmoe.moe_dir is not None, "MoE weights should be provided"
        self.moe = MixtureOfExperts(params)
        self.moe.load_state_dict(torch.load(params.pmoe.moe_dir, map_location="cpu"), strict=False)
        self.moe = freeze(self.moe, params.exclude_freeze, params.verbose)
        self.punet = PUNetExpert(params)
        if params.pmoe.punet_dir:
            self.punet.load_state_dict(torch.load(params.pmoe.punet_dir, map_location="cpu"), strict=False)
            self.punet = freeze(self.punet, params.exclude_freeze, params.verbose)
        self.lat_weights = nn.Linear(2, 1)
        self.long_weights = nn.Linear(2, 1)

    def forward(self, images, speed, command):
        punet_actions, _ = self.punet(images.clone(), speed.clone(), command.clone())
        probs, mean, std, _, _ = self.moe.components(images, speed, command)
        moe_actions = draw(probs, mean, std)  # dists.sample() of the reference: no gradient into the mixture
        # the 2->1 combiners are 8 FLOPs per sample: left to ATen on the GPU
        lat = self.lat_weights(torch.cat([moe_actions[:, 0:1], punet_actions[:, 0:1]], dim=-1))
        lon = self.long_weights(torch.cat([moe_actions[:, 1:], punet_actions[:, 1:]], dim=-1))
        return torch.tanh(torch.cat([lat, lon], dim=-1)), -1

    def sample(self, images, speed, command) -> torch.Tensor:
        actions, _ = self.forward(images, speed, command)
        return actions
