"""Golden vectors for the GPU input pipeline: the reference's eval-mode frame transform run with the REAL third-party
calls it makes — Crop (PMoE/model/augmenter.py:43-49: Image.fromarray(img[top:-bottom])), torchvision
transforms.Resize(resize) on the PIL image, transforms.ToTensor() (PMoE/model/data_loader.py:275-281) — on small seeded
uint8 frames. Writes tests/golden/preproc.pt. Run here (torchvision + Pillow are importable in this container)."""
import os

import numpy as np
import torch
from PIL import Image
import PIL
import torchvision
from torchvision import transforms

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def reference_transform(img, crop, resize):
    pil = Image.fromarray(img[crop[0]:-crop[1]])                       # augmenter.py:48-49
    return transforms.Compose([transforms.Resize(resize), transforms.ToTensor()])(pil)   # data_loader.py:275-281


def main():
    g = np.random.default_rng(20240)
    cases = []
    for (hs, ws, crop, resize) in [(240, 320, (125, 90), (64, 80)), (271, 333, (50, 31), (224, 224)), (96, 120, (10, 6), (112, 160)),
                                   (260, 224, (20, 16), (224, 224)), (420, 560, (125, 90), (224, 224))]:
        # smooth content + noise + saturated patches: exercises rounding, clipping and both up- and down-scaling
        base = g.integers(0, 256, size=(2 if hs <= 240 else 1, hs // 7 + 2, ws // 7 + 2, 3), dtype=np.uint8)
        frames = np.stack([np.asarray(Image.fromarray(b).resize((ws, hs), Image.BICUBIC)) for b in base])
        frames = np.clip(frames.astype(np.int32) + g.integers(-40, 41, size=frames.shape), 0, 255).astype(np.uint8)
        frames[:, hs // 2:hs // 2 + 9, :30] = 255
        frames[:, hs // 2 + 9:hs // 2 + 15, :30] = 0
        out = torch.stack([reference_transform(f, crop, resize) for f in frames])
        # ToTensor is uint8 -> float / 255: the fixture keeps the uint8 values (4x smaller); the test redoes the division
        out_u8 = (out * 255).round().to(torch.uint8)
        assert torch.equal(out_u8.float().div(255), out)
        cases.append({"frames": torch.from_numpy(frames), "crop": crop, "resize": resize, "out_u8": out_u8})
    torch.save({"cases": cases, "pillow": PIL.__version__, "torchvision": torchvision.__version__},
               os.path.join(ROOT, "tests", "golden", "preproc.pt"))
    print("wrote", len(cases), "cases; Pillow", PIL.__version__, "torchvision", torchvision.__version__)


if __name__ == "__main__":
    main()
