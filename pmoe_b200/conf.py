"""The `model:` sub-tree of the reference's conf/stage_2*.yaml as a plain attribute-dict, for callers that do not
carry OmegaConf (bench.py, tests). Values are the shipped defaults (PMoE/conf/stage_2.yaml:76-133); `pretrained` is
False because there is no network for ImageNet weights, and `device` is the key PUNetExpert reads (moe.py:278)."""
import copy


class AttrDict(dict):
    """dict with attribute access that also supports ** splatting and item assignment, like a DictConfig."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v

    def __deepcopy__(self, memo):
        return AttrDict({k: copy.deepcopy(v, memo) for k, v in self.items()})


def wrap(o):
    if isinstance(o, dict):
        return AttrDict({k: wrap(v) for k, v in o.items()})
    if isinstance(o, (list, tuple)):
        return [wrap(v) for v in o]
    return o


def _mlp(dims, act, l_act, dropout):
    return {"dims": list(dims), "act": act, "l_act": l_act, "bn": False, "dropout": dropout}


def stage2_model_cfg(type="moe_alt", n_experts=3, n_frames=4, dropout=0.3, device="cuda"):
    return wrap({
        "verbose": False, "type": type, "n_experts": n_experts, "loss_coefs": [0.7, 0.3], "exclude_freeze": [],
        "punet_path": "", "device": device,
        "action_head": _mlp([1536, 512, 512], "elu", True, dropout),
        "speed_encoder": _mlp([1, 512, 512], "relu", False, dropout),
        "command_encoder": _mlp([6, 512, 512], "relu", False, dropout),
        "speed_prediction": _mlp([1536, 512, 512, 1], "relu", False, dropout),
        "backbone": {"type": "rgb", "n_frames": n_frames,
                     "rgb": {"arch": "resnet18", "pretrained": False, "gamma": 2, "b": 1},
                     "segmentation": {"gamma": 2, "b": 1, "inter_repr": True, "model_dir": ""}},
        "punet": {"past_frames": 4, "future_frames": 6, "in_features": 3, "num_classes": 23, "gamma": 2, "b": 1,
                  "unet_inter_repr": False, "model_name": "unet", "model_path": ""},
        "pmoe": {"moe_dir": "", "punet_dir": ""},
    })
