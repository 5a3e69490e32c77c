"""Behavior_policy.learn: the register-tiled kernels (csrc/beh_learn_tile.cu) against the one-warp-per-chain draft
(csrc/beh_learn.cu, itself pinned to the reference's recorded call) on the same inputs and dropout masks, at a shape the
golden fixtures do not reach (several 64-chain tiles, a ragged last tile, terminated agents), and the time of one call.

    timeout 600 python tools/check_beh_learn_tile.py [--envs 6] [--steps 25] [--time-envs 512]
"""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from iplan_b200 import _lib                                              # noqa: E402
from iplan_b200.components.episode_buffer import EpisodeBatch            # noqa: E402
from iplan_b200.config import make_args                                  # noqa: E402
from iplan_b200.nova.stable_behavior_policy import Behavior_policy       # noqa: E402
from tools.check_pred_learn import scheme_for                            # noqa: E402


def make_batch(args, B, T1, seed):
    A, N, o = args.n_agents, args.max_vehicle_num, args.obs_shape_single
    g = torch.Generator().manual_seed(seed)
    scheme, groups, pre = scheme_for(args)
    batch = EpisodeBatch(scheme, groups, B, T1, preprocess=pre, device="cuda")
    hist = torch.rand(B, T1, A, N, o, generator=g) * 2 - 1
    hist[..., 0] = 1.0
    for t in range(T1):
        hist[:, t, :, min(N, 15 + t // 3):] = 0.0
    term = (torch.cumsum((torch.rand(B, T1, A, generator=g) < 0.08).int(), dim=1) > 0).to(torch.uint8).unsqueeze(-1)
    batch.update({"history": hist.numpy(), "terminated": term.numpy()}, bs=slice(None), ts=slice(None))
    return batch


def run(envs=6, steps=25, seed=3):
    args = make_args("highway", use_cuda=True, device="cuda")
    A, N, W = args.n_agents, args.max_vehicle_num, args.max_history_len
    T1 = steps + 1
    n_pos = T1 - 1 - 1 - W
    batch = make_batch(args, envs, T1, seed)
    keep = (torch.rand(A, envs, n_pos, N, W, args.decoder_rnn_dim, generator=torch.Generator().manual_seed(seed + 1)) >= args.decoder_dropout).to(torch.uint8)
    out = {}
    for impl in (1, 0):
        torch.manual_seed(seed)
        pol = Behavior_policy(args, None)
        _lib.lib.iplan_beh_learn_set_impl(impl)
        pol.debug_keep = keep
        b_loss, s_loss, _ = pol.learn(batch, t_env=0)
        torch.cuda.synchronize()
        out[impl] = dict(b=[float(x) for x in b_loss], s=[float(x) for x in s_loss], enc=pol.last_grads["enc"].cpu(), dec=pol.last_grads["dec"].cpu(),
                         names=(pol.stack.named_offsets(), pol.dec_stack.named_offsets()))
    _lib.lib.iplan_beh_learn_set_impl(0)
    ok = True
    for a in range(A):
        db = abs(out[0]["b"][a] - out[1]["b"][a]) / abs(out[1]["b"][a])
        ds = abs(out[0]["s"][a] - out[1]["s"][a]) / max(abs(out[1]["s"][a]), 1e-12)
        print(f"[B={envs} T={steps} a={a}] behavior loss tile {out[0]['b'][a]:.6f} draft {out[1]['b'][a]:.6f} (rel {db:.2e}); stability rel {ds:.2e}")
        ok &= db < 1e-5 and ds < 1e-5
        for kind, names in (("enc", out[0]["names"][0]), ("dec", out[0]["names"][1])):
            for name, (off, shape) in names.items():
                n = max(1, int(torch.tensor(shape).prod()))
                mine, want = out[0][kind][a, off:off + n], out[1][kind][a, off:off + n]
                rel = float((mine - want).abs().max() / (want.abs().max() + 1e-12))
                flagged = "" if rel < 1e-4 else "   <-- MISMATCH"
                if a == 0 or flagged:
                    print(f"    grad {kind}:{name:28s} |draft| {float(want.abs().max()):.3e} rel {rel:.2e}{flagged}")
                ok &= rel < 1e-4
    return ok


def timing(envs, steps=90):
    args = make_args("highway", use_cuda=True, device="cuda")
    batch = make_batch(args, envs, steps + 1, 5)
    pol = Behavior_policy(args, None)
    _lib.lib.iplan_beh_learn_set_impl(0)
    for it in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        pol.learn(batch, t_env=0)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f"Behavior_policy.learn (tile kernels) at {envs} envs, T={steps}: {dt * 1e3:.1f} ms (call {it})")
    return dt


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=6)
    ap.add_argument("--steps", type=int, default=25)
    ap.add_argument("--time-envs", type=int, default=0)
    a = ap.parse_args()
    good = run(a.envs, a.steps)
    print("OK" if good else "MISMATCH")
    if a.time_envs:
        timing(a.time_envs)
    sys.exit(0 if good else 1)
