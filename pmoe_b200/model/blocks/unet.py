"""U-Net with the reference's constructor and state_dict (reference: PMoE/model/blocks/unet.py:8-95)."""
import torch
import torch.nn as nn

from ... import config, infer, nhwc, train
from .basics import EfficientBlock, conv3, _grad_mode


class UNet(nn.Module):
    def __init__(self, in_features: int = 3, out_features: int = 23, gamma: int = 2, b: int = 1, dropout: float = 0.0,
                 inter_repr: bool = False):
        super().__init__()
        if dropout != 0.0:
            raise NotImplementedError("pmoe_b200 UNet: Dropout2d p>0 is not used by any reference config (unet.py:17,31)")
        self.inter_repr = inter_repr
        widths = ((in_features, 64), (64, 128), (128, 256), (256, 512), (512, 512))
        for i, (ci, co) in enumerate(widths, start=1):
            setattr(self, "dwn_%d" % i, conv3(ci, co))
        self.pool = nn.MaxPool2d(kernel_size=2, stride=2)
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))
        self.dropout = nn.Dropout2d(p=dropout)
        for i, (ci, co) in enumerate(((512, 512), (512, 256), (256, 128), (128, 64)), start=1):
            setattr(self, "up_%d" % i, nn.ConvTranspose2d(ci, co, kernel_size=2, stride=2))
            setattr(self, "up_forw_%d" % i, conv3(2 * co, co))
        self.out = nn.Conv2d(64, out_features, kernel_size=1)

    def forward(self, image):
        """image: fp32 (B,C,H,W) on the GPU -> logits (B,classes,H,W) [, after (B,512) if inter_repr]."""
        if _grad_mode(self):
            return train.unet_module_forward(self, image)
        x = nhwc.from_nchw(image, dtype=config.act_dtype())
        n, _, h, w = image.shape
        out = torch.empty(n, self.out.weight.shape[0], h, w, dtype=torch.float32, device=image.device)
        _, inter = infer.unet_eval(self, x, want_inter=self.inter_repr, nchw_out=out)
        if self.inter_repr:
            return inter, out
        return out


class UNetECA(nn.Module):
    """U-Net with ECA coefficients (reference: PMoE/model/blocks/unet.py:98-185): 32..512 channels, an ECA gate on the
    pooled x_4 and on every decoder concatenation. No reference trainer or config instantiates it; it runs on the general
    tape executor in both modes (BatchNorm follows `.training`), not on the specialised eval path."""

    def __init__(self, in_features: int = 3, out_features: int = 23, gamma: int = 2, b: int = 1, dropout: float = 0.0,
                 inter_repr: bool = False):
        super().__init__()
        if dropout != 0.0:
            raise NotImplementedError("pmoe_b200 UNetECA: Dropout2d p>0 is not used by any reference config (unet.py:108,126)")
        self.inter_repr = inter_repr
        for i, (ci, co) in enumerate(((in_features, 32), (32, 64), (64, 128), (128, 256), (256, 512)), start=1):
            setattr(self, "dwn_%d" % i, conv3(ci, co))
        self.eca_0 = EfficientBlock(512, gamma, b)
        self.pool = nn.MaxPool2d(kernel_size=2, stride=2)
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))
        self.dropout = nn.Dropout2d(p=dropout)
        for i, (ci, co) in enumerate(((512, 256), (256, 128), (128, 64), (64, 32)), start=1):
            setattr(self, "up_%d" % i, nn.ConvTranspose2d(ci, co, kernel_size=2, stride=2))
            setattr(self, "eca_%d" % i, EfficientBlock(2 * co, gamma, b))
            setattr(self, "up_forw_%d" % i, conv3(2 * co, co))
        self.out = nn.Conv2d(32, out_features, kernel_size=1)

    def forward(self, image):
        """image: fp32 (B,C,H,W) on the GPU -> logits (B,classes,H,W) [, after the pooled dwn_5 output (B,512) if inter_repr]."""
        return train.unet_module_forward(self, image, body=train.unet_eca)
