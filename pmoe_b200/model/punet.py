"""Predictive U-Net (reference: PMoE/model/punet.py:12-120): same constructor arguments, checkpoint
side effects (loads `model_path`[`model_name`] into the frozen U-Net) and state_dict keys."""
import torch
import torch.nn as nn

from .. import infer, train
from .blocks.basics import EfficientConvBlock, _grad_mode
from .blocks.unet import UNet


class PredictiveUnet(nn.Module):
    def __init__(self, past_frames: int = 4, future_frames: int = 4, in_features: int = 3, num_classes: int = 23,
                 gamma: int = 2, b: int = 1, inter_repr: bool = False, unet_inter_repr: bool = False,
                 model_name: str = "unet-swa", model_path: str = "unet.pth"):
        super().__init__()
        self.n_past_frames = past_frames
        self.n_future_frames = future_frames
        self.inter_repr = inter_repr
        self.unet_inter_repr = unet_inter_repr
        self.unet = UNet(in_features=in_features, out_features=num_classes, gamma=gamma, b=b, inter_repr=unet_inter_repr)
        checkpoint = torch.load(model_path, map_location="cpu")
        self.unet.load_state_dict(checkpoint[model_name], strict=False)
        for p in self.unet.parameters():
            p.requires_grad = False
        self.unet.eval()
        self.entry_block = EfficientConvBlock(in_ch=past_frames * num_classes, out_ch=in_features, gamma=gamma, b=b)
        self.pred_unet = UNet(in_features=in_features, out_features=num_classes, gamma=gamma, b=b, inter_repr=inter_repr)

    def forward(self, img_list: torch.Tensor) -> torch.Tensor:
        assert img_list.shape[-4] == self.n_past_frames, "Number of images should match number of past frames"
        if _grad_mode(self):
            return train.punet_module_forward(self, img_list)
        return infer.punet_eval(self, img_list)
