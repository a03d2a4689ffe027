"""One training step of the ResNet18-ECA STEM alone (EfficientConvBlock 12->64->64 at full resolution, torchvision's bn1 + ReLU,
MaxPool2d(3,2,1)) — the part of an expert where the large memory-bound launches live (822 MB tensors at B=128). Small launch
count, so that an `ncu --set full` pass over every kernel stays short:  python scripts/gpu_stem_step.py [batch] [steps]"""
import os
import sys

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pmoe_b200 import train
from pmoe_b200.model.blocks.basics import EfficientConvBlock

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2


class Stem(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv1 = EfficientConvBlock(12, 64)
        self.bn1 = nn.BatchNorm2d(64)

    def forward(self, x):
        def body(tape, a):
            y = train.eca_conv_block(tape, self.conv1, a, tag="stem.conv1", want_out_stats=True)
            y = train.bn_act_op(tape, self.bn1, y, "relu", tag="stem.bn1")
            return train.maxpool_op(tape, y, 3, 2, 1)
        return train.nhwc_module_forward(self, x, body)


torch.manual_seed(0)
m = Stem().cuda().train()
x = torch.rand(B, 12, 224, 224, device="cuda")
for _ in range(steps):
    m.zero_grad()
    out = m(x)
    out.backward(torch.ones_like(out) * 1e-3)
torch.cuda.synchronize()
print("stem step ok", tuple(out.shape), float(m.bn1.weight.grad.norm()))
