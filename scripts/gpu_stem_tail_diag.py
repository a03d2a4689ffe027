"""Diagnostic: run-to-run floor of a ResNet-18 bf16 step vs the fused/unfused stem tail difference (forward + gradients)."""
import json, sys, torch
from pmoe_b200 import config, train
from pmoe_b200.model.blocks.backbone import get_backbone
dev = "cuda"
def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()
res = {}
for (B, HW) in ((4, 64), (8, 128)):
    torch.manual_seed(3)
    x = torch.rand(B, 12, HW, HW, device=dev)
    cot = torch.randn(B, 512, device=dev)
    caps = []
    orig_fused, orig_pool = train.bn_relu_maxpool_op, train.maxpool_op
    def cap_fused(tape, bn, xx, tag=""):
        pa = orig_fused(tape, bn, xx, tag=tag)
        caps.append((xx.t.detach().clone(), pa.t.detach().clone()))
        return pa
    train.bn_relu_maxpool_op = cap_fused
    with config.use_precision("bf16"):
        net = get_backbone(arch="resnet18", n_frames=4, pretrained=False, gamma=2, b=1, n_channels=3).cuda().train()
        sd = {k: v.clone() for k, v in net.state_dict().items()}
        runs = []
        for fused in (True, False, False, True):
            net.load_state_dict(sd); net.zero_grad(set_to_none=True)
            train.FUSE_STEM_TAIL = fused
            f = net(x); (f * cot).sum().backward()
            runs.append((f.detach().clone(), {n: p.grad.detach().clone() for n, p in net.named_parameters()}))
    train.bn_relu_maxpool_op = orig_fused
    def gerr(a, b):
        num = sum((a[n].double() - g.double()).pow(2).sum() for n, g in b.items()); den = sum(g.double().pow(2).sum() for g in b.values())
        return (num / den).sqrt().item()
    r = {"fwd_fused_vs_unfused": rel(runs[0][0], runs[1][0]), "fwd_unfused_vs_unfused": rel(runs[2][0], runs[1][0]),
         "fwd_fused_vs_fused": rel(runs[3][0], runs[0][0]),
         "grad_fused_vs_unfused": gerr(runs[0][1], runs[1][1]), "grad_unfused_vs_unfused": gerr(runs[2][1], runs[1][1]),
         "grad_fused_vs_fused": gerr(runs[3][1], runs[0][1]),
         "stem_equal": [bool(torch.equal(caps[0][0], c[0])) for c in caps], "pool_equal": [bool(torch.equal(caps[0][1], c[1])) for c in caps],
         "pool_rel": [rel(c[1].float(), caps[0][1].float()) for c in caps], "stem_rel": [rel(c[0].float(), caps[0][0].float()) for c in caps]}
    res["B%d_%d" % (B, HW)] = r
    print(json.dumps(r, indent=1))
json.dump(res, open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/stem_tail_diag.json", "w"), indent=1)
