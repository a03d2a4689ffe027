"""BASELINE configs[3]: MoE head isolation — gating + grouped expert MLPs on 65536 feature vectors, K in {4, 8, 16}.
Per expert (conf/stage_2.yaml:83-106): speed_pred [1536,512,512,1] ReLU, action_features [1536,512,512] ELU(+last),
action_pred 512->4, alpha 512->1 (+ReLU), then softmax_K, ELU+1, mixture NLL + speed MSE; dropout off.
Forward + backward through the tape (grouped GEMM over the expert axis) and, for K=4, parity of the forward against a
plain torch fp32 evaluation of the same heads on the bf16-rounded operands. Writes gpurun_out/heads_bench.json."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pmoe_b200 import conf, config, loss as L, profiler, train
from pmoe_b200.model.blocks.basics import make_mlp
from pmoe_b200.model.moe import _mixture

FLOPS_PER_VEC_EXPERT = 2.0 * (1536 * 512 + 512 * 512 + 512 * 1 + 1536 * 512 + 512 * 512 + 512 * 4 + 512 * 1)  # 4.201 MF


class Heads(torch.nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.speed_pred = make_mlp(**cfg.speed_prediction)
        self.action_features = make_mlp(**cfg.action_head)
        self.action_pred = torch.nn.Linear(512, 4)
        self.alpha = torch.nn.Linear(512, 1)


class HeadBank(torch.nn.Module):
    def __init__(self, K):
        super().__init__()
        cfg = conf.stage2_model_cfg("moe", K, dropout=0.0)
        self.K = K
        self.moe = torch.nn.ModuleList([Heads(cfg) for _ in range(K)])

    def forward(self, feats):  # feats: (K, 1, B, 1536) bf16 on the GPU
        K, B = self.K, feats.shape[2]

        def runner(tape):
            f = train._new_act(tape, feats, 1536, False)
            ex = list(self.moe)
            sp = train.grouped_mlp(tape, [e.speed_pred for e in ex], [f], "speed_pred")
            af = train.grouped_mlp(tape, [e.action_features for e in ex], [f], "action_features")
            ap = train.grouped_linear_op(tape, [af], [e.action_pred for e in ex], None, tag="action_pred")
            al = train.grouped_linear_op(tape, [af], [e.alpha for e in ex], None, tag="alpha")
            gm = train.GateMixture(tape, [al], [ap], al.t, ap.t, B, K, relu_alpha=True, a_sk=B * 16, p_sk=B * 16)
            speeds = sp.t.view(K, B, 16)[:, :, :1].permute(1, 0, 2).float().contiguous()

            def seed(tp, g):
                gm.backward(g[0], g[1], g[2])
                if g[3] is not None:
                    train.seed_stacked(tp, sp, g[3])
            return [gm.probs, gm.mean, gm.std, speeds, gm.route], seed

        return train.run(self, runner)


def torch_reference(bank, feats):
    """fp32 torch evaluation of the same heads on the bf16-rounded weights/inputs."""
    outs = []
    for e, h in enumerate(bank.moe):
        x = feats[e, 0].float()
        r = lambda lin, v: torch.nn.functional.linear(v, lin.weight.to(torch.bfloat16).float(), lin.bias.float())
        sp_l = [m for m in h.speed_pred if isinstance(m, torch.nn.Linear)]
        af_l = [m for m in h.action_features if isinstance(m, torch.nn.Linear)]
        s = torch.relu(r(sp_l[0], x))
        s = torch.relu(r(sp_l[1], s.to(torch.bfloat16).float()))
        s = r(sp_l[2], s.to(torch.bfloat16).float())
        a = torch.nn.functional.elu(r(af_l[0], x))
        a = torch.nn.functional.elu(r(af_l[1], a.to(torch.bfloat16).float())).to(torch.bfloat16).float()
        ap = r(h.action_pred, a)
        al = torch.relu(r(h.alpha, a))
        outs.append((al, ap, s))
    alpha = torch.cat([o[0] for o in outs], 1)
    probs = torch.softmax(alpha.to(torch.bfloat16).float().clamp_min(0), 1)
    return probs, alpha


def main():
    config.set_precision("bf16")
    B = int(os.environ.get("HEADS_B", "65536"))
    report = []
    # The first graph capture of a process fails on lazy initialisation (both attempts, whatever K comes first); a throw-away
    # small case absorbs it so that every reported K is timed as a graph replay.
    ks = [int(k) for k in os.environ.get("HEADS_KS", "4,8,16").split(",")]
    for K in [-2] + ks:
        throwaway = K < 0
        K = abs(K)
        B = 4096 if throwaway else int(os.environ.get("HEADS_B", "65536"))
        torch.manual_seed(K)
        bank = HeadBank(K).cuda().train()
        g = torch.Generator().manual_seed(1)
        feats = torch.randn(K, 1, B, 1536, generator=g).to(torch.bfloat16).cuda()
        control = (torch.rand(B, 2, generator=g) * 2 - 1).cuda()
        target = torch.rand(B, 1, generator=g).cuda()

        def step():
            for p in bank.parameters():
                p.grad = None
            probs, mean, std, speeds, route = bank(feats)
            loss = L.moe_loss(_mixture(probs, mean, std), speeds, control, target.clone(), [0.7, 0.3])
            loss.backward()
            return probs, route, loss

        for _ in range(3):  # the first step packs weights and grows the allocator pool (hundreds of ms)
            probs, route, loss = step()
        torch.cuda.synchronize()
        # The eager step is dominated by the caching allocator at this size (GB-scale activations freed and re-requested every
        # step: 11..100 ms of host time around ~7 ms of kernels, depending on what the caller keeps alive), so the timed
        # region replays a CUDA graph of forward + loss + backward, like the training leg of bench.py.
        mode, run, launches = None, None, None
        torch.distributions.Distribution.set_default_validate_args(False)
        for attempt in range(2):  # the first capture of a process can trip over lazy initialisation; the second one is clean
            try:
                for p in bank.parameters():
                    if p.grad is not None:
                        p.grad.zero_()
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    probs, mean_, std_, speeds_, route = bank(feats)
                    L.moe_loss(_mixture(probs, mean_, std_), speeds_, control, target.clone(), [0.7, 0.3]).backward()
                torch.cuda.current_stream().wait_stream(side)
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                profiler.reset()
                with torch.cuda.graph(graph):
                    probs, mean_, std_, speeds_, route = bank(feats)
                    loss = L.moe_loss(_mixture(probs, mean_, std_), speeds_, control, target.clone(), [0.7, 0.3])
                    loss.backward()
                launches = profiler.launch_count()
                mode, run = "cuda_graph", graph.replay
                break
            except Exception as ex:
                print("capture attempt %d failed: %s" % (attempt, str(ex)[:600]), flush=True)
                mode = "eager (%s)" % str(ex).splitlines()[0][:80]
                torch.cuda.synchronize()
        if run is None:
            def run():
                global_out[:] = step()
            global_out = [None, None, None]
            run()
            probs, route, loss = global_out
        for _ in range(2):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 10
        if launches is None:
            profiler.reset()
        e0.record()
        for _ in range(n):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        if launches is None:
            launches = profiler.launch_count() // n
        rec = {"K": K, "B": B, "ms_fwd_bwd": ms, "vectors_per_s": B / ms * 1e3, "tflops": 3 * FLOPS_PER_VEC_EXPERT * K * B / ms / 1e9,
               "launches_per_step": launches, "loss": float(loss.item()), "launch_mode": mode}
        if K == 4:
            with torch.no_grad():
                pr, alpha = torch_reference(bank, feats)
            rec["probs_rel_err_vs_torch"] = ((probs - pr).norm() / pr.norm()).item()
            # routing index: compare where the torch logits are not within bf16 rounding of a tie
            top2 = alpha.topk(2, dim=1).values
            clear = (top2[:, 0] - top2[:, 1]) > 2e-2 * top2[:, 0].abs().clamp_min(1e-3)
            rec["route_agree_clear"] = float((route[clear] == alpha.argmax(1)[clear]).float().mean().item())
            rec["route_clear_fraction"] = float(clear.float().mean().item())
        if throwaway:
            del bank, feats
            torch.cuda.empty_cache()
            continue
        print(json.dumps(rec), flush=True)
        if os.environ.get("HEADS_PROFILE"):
            profiler.enable_events(True)
            step()
            torch.cuda.synchronize()
            rows = sorted(((ms_, k_, t_) for (k_, ms_, f_, b_, t_) in profiler.records()), reverse=True)
            for ms_, k_, t_ in rows[:14]:
                print("   %8.3f ms  %-16s %s" % (ms_, k_, t_))
            print("   sum of profiled launches %.2f ms over %d" % (sum(r[0] for r in rows), len(rows)))
            import collections
            cnt, tms = collections.Counter(), collections.Counter()
            for ms_, k_, t_ in rows:
                cnt[k_] += 1
                tms[k_] += ms_
            print("   launches by kernel: " + ", ".join("%s x%d (%.2f ms)" % (k_, n_, tms[k_]) for k_, n_ in cnt.most_common()))
            rec["launches_by_kernel"] = {k_: [n_, tms[k_]] for k_, n_ in cnt.most_common()}
            profiler.enable_events(False)
        report.append(rec)
        del bank, feats
        torch.cuda.empty_cache()
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(report, open("gpurun_out/heads_bench.json", "w"), indent=1)


if __name__ == "__main__":
    main()
