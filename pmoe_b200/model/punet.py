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

    def forward(self, img_list: torch.Tensor, host_out: torch.Tensor = None, copy_stream=None) -> torch.Tensor:
        """Reference signature `forward(img_list)` (punet.py:75). Serving extension, eval mode only: with `host_out` (a
        pinned fp32 host tensor from `infer.pinned_output_like`, shape (B, F, classes, H, W)) every future frame's logits
        are streamed to the host on `copy_stream` as soon as the U-Net pass that produced them has finished, overlapping
        the later passes; the caller synchronises `copy_stream` before reading `host_out`. The device tensor is returned
        either way."""
        assert img_list.shape[-4] == self.n_past_frames, "Number of images should match number of past frames"
        if _grad_mode(self):
            if host_out is not None:
                raise RuntimeError("pmoe_b200 PredictiveUnet: host_out streaming is an eval-mode (no_grad) feature")
            return train.punet_module_forward(self, img_list)
        return infer.punet_eval(self, img_list, host_out=host_out, copy_stream=copy_stream)
