"""Behavior_policy.learn on the GPU (kernel csrc/beh_learn.cu) against the reference's recorded ``learn`` call
(tests/golden/behavior_learn.pt, dropout masks replayed):

    timeout 200 python tools/check_beh_learn.py

Prints, per case and tensor, the relative difference of the raw gradients to the oracle's (autograd) and of the post-step
weights to the reference's, so that a partial failure localises the faulty phase of the kernel."""
import os
import sys
from types import SimpleNamespace

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from iplan_b200.components.episode_buffer import EpisodeBatch            # noqa: E402
from iplan_b200.config import make_args                                  # noqa: E402
from iplan_b200.nova.stable_behavior_policy import Behavior_policy       # noqa: E402
from oracle import iplan_oracle as O                                     # noqa: E402
from tools.check_pred_learn import scheme_for                            # noqa: E402


def run(case):
    g = torch.load(os.path.join(ROOT, "tests", "golden", "behavior_learn.pt"), weights_only=False)[case]
    args = make_args(g["args"].get("env", "highway"))
    for k, v in g["args"].items():
        setattr(args, k, v)
    args.use_cuda, args.device = True, "cuda"
    A, N, W = args.n_agents, args.max_vehicle_num, args.max_history_len
    d = g["data"]
    B, T1 = d["history"].shape[:2]
    n_pos = T1 - 1 - 1 - W
    scheme, groups, pre = scheme_for(args)
    batch = EpisodeBatch(scheme, groups, B, T1, preprocess=pre, device="cuda")
    batch.update({k: v.numpy() for k, v in d.items()}, bs=slice(None), ts=slice(None))
    pol = Behavior_policy(args, None)
    for a in range(A):
        pol.behavior_encoder[a].load_state_dict(g["enc_before"][a])
        pol.behavior_decoder[a].load_state_dict(g["dec_before"][a])
    pol.debug_keep = torch.stack([k.view(n_pos, B, N, W, -1).permute(1, 0, 2, 3, 4) for k in g["dropout_keep"]]).to(torch.uint8)
    b_loss, s_loss, _ = pol.learn(batch, t_env=0)
    torch.cuda.synchronize()
    ok = True
    hist = d["history"][:, :-1]
    term = d["terminated"][:, :-1, :, 0].float()
    oargs = SimpleNamespace(**g["args"])
    for a in range(A):
        mask = 1 - term[:, :, a] if oargs.env == "MPE" else term[:, :, a]
        ep = {k: v.clone() for k, v in g["enc_before"][a].items()}
        dp = {k: v.clone() for k, v in g["dec_before"][a].items()}
        # raw (unclipped) gradients from the oracle: rerun its graph without the clip
        ref, _ = O.behavior_learn_agent(ep, dp, hist[:, :, a], mask, g["dropout_keep"][a], oargs)
        db = abs(float(b_loss[a]) - g["behavior_loss"][a]) / abs(g["behavior_loss"][a])
        ds = abs(float(s_loss[a]) - g["stability_loss"][a]) / abs(g["stability_loss"][a])
        print(f"[{case} a={a}] behavior loss cuda {float(b_loss[a]):.6f} reference {g['behavior_loss'][a]:.6f} (rel {db:.2e}); "
              f"stability {float(s_loss[a]):.6f} vs {g['stability_loss'][a]:.6f} (rel {ds:.2e})")
        ok &= db < 1e-4 and ds < 1e-4
        # the oracle returns the gradients as clipped by the reference; clip ours the same way per group
        for kind, stack, flat in (("enc", pol.stack, pol.last_grads["enc"]), ("dec", pol.dec_stack, pol.last_grads["dec"])):
            mine_all = {name: flat[a, off:off + max(1, int(torch.tensor(shape).prod()))].view(shape).cpu()
                        for name, (off, shape) in stack.named_offsets().items()}
            total = torch.sqrt(sum((v ** 2).sum() for v in mine_all.values()))
            coef = min(1.0, float(args.max_grad_norm) / (float(total) + 1e-6))
            for name, mine in mine_all.items():
                want = ref["clipped"][kind + ":" + name]
                rel = float((mine * coef - want).abs().max() / (want.abs().max() + 1e-12))
                flagged = "" if rel < 1e-3 else "   <-- MISMATCH"
                print(f"    grad {kind}:{name:28s} rel {rel:.2e}{flagged}")
                ok &= rel < 1e-3
        after = {**{"enc:" + k: v for k, v in pol.behavior_encoder[a].state_dict().items()},
                 **{"dec:" + k: v for k, v in pol.behavior_decoder[a].state_dict().items()}}
        want = {**{"enc:" + k: v for k, v in g["enc_after"][a].items()}, **{"dec:" + k: v for k, v in g["dec_after"][a].items()}}
        worst = max(float((after[k].cpu() - want[k]).abs().max()) for k in want)
        print(f"    max |weight - reference| after the step: {worst:.2e}")
        ok &= worst < 1e-6
    return ok


if __name__ == "__main__":
    res = [run(c) for c in ("mpe", "highway")]
    print("OK" if all(res) else "MISMATCH")
