"""One conv shape, a few launches — the target of `ncu --set full`."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pmoe_b200 import ops
B, H, cin, cout = [int(x) for x in sys.argv[1:5]] if len(sys.argv) >= 5 else (32, 224, 64, 64)
dev = "cuda"
x = torch.randn(B, H, H, ops.pad_ch(cin), device=dev).to(torch.bfloat16)
w = (torch.randn(cout, cin, 3, 3, device=dev) / (cin * 9) ** 0.5)
cp = [ops.pad_ch(cin)]
ck = ops.choose_ck(cp)
cop = ops.cout_padded(cout)
wp = ops.pack_conv_weight(w, [cin], cp, ops.TAPS3, cop)
segs = ops.conv_segments([(r - 1, s - 1) for (r, s) in ops.TAPS3], cp, ck)
out = torch.empty(B, H, H, ops.pad_ch(cout), dtype=torch.bfloat16, device=dev)
sc = None  # eval-mode BN scale lives in the packed weights: the same specialised epilogue (shift + ReLU) the model runs
sh = torch.zeros(cop, device=dev)
for _ in range(3):
    ops.conv_tc([x], wp, segs, ck, out, sc, sh, "relu")
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    ops.conv_tc([x], wp, segs, ck, out, sc, sh, "relu")
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(json.dumps({"B": B, "H": H, "cin": cin, "cout": cout, "ms": ms, "tflops": 2.0 * B * H * H * cin * cout * 9 / ms / 1e9}))
