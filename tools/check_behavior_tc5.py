"""K1b: the tcgen05 kernel (csrc/behavior_tc5.cu, impl 0) against the mma.sync kernel (csrc/behavior_step.cu, impl 1) on the
same inputs — contiguous windows (the numpy API's layout) and windows read in place from a time-strided store with leading
zero rows (the device-resident runner's layout), ragged chain counts — and the time of both at the bench shape.

    timeout 300 python tools/check_behavior_tc5.py [--time-envs 512]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from iplan_b200 import _lib                                              # noqa: E402
from iplan_b200.config import make_args                                  # noqa: E402
from iplan_b200.nova.stable_behavior_policy import Behavior_policy       # noqa: E402


def run_pair(pol, B, mode, seed):
    args = pol.args
    A, N, o, L, W, E = args.n_agents, args.max_vehicle_num, args.obs_shape_single, args.latent_dim, args.max_history_len, args.encoder_rnn_dim
    g = torch.Generator(device="cuda").manual_seed(seed)
    hid0 = torch.rand(A, B, N, E, device="cuda", generator=g) * 2 - 1
    prev = torch.softmax(torch.randn(A, B, N, L, device="cuda", generator=g), dim=-1)
    out = {}
    if mode == "contiguous":
        win = torch.rand(A, B, N, W * o, device="cuda", generator=g) * 2 - 1
        kw = dict()
        window = win
    else:                                   # rows of a [A,B,T,N,S] store, S = 45 floats per slot, window = times first .. first + W - 1 - pad
        T, S, pad = 14, 45, 3 if mode == "strided_pad" else 0
        store = torch.rand(A, B, T, N, S, device="cuda", generator=g) * 2 - 1
        first = 2
        window = store[:, :, first, :, :o]
        kw = dict(win_stride_step=store.stride(2), win_pad=pad)
    for impl in (1, 0):
        _lib.lib.iplan_behavior_set_impl(impl)
        hid = hid0.clone()
        new = torch.empty(A, B, N, L, device="cuda")
        pol.behavior_step(window, hid, prev, new, **kw)
        torch.cuda.synchronize()
        out[impl] = (new.cpu(), hid.cpu())
    d_lat = float((out[0][0] - out[1][0]).abs().max())
    d_hid = float((out[0][1] - out[1][1]).abs().max())
    print(f"[B={B} {mode}] |latent tc5 - mma.sync| {d_lat:.2e}   |hidden| {d_hid:.2e}   (|hidden| max {float(out[1][1].abs().max()):.2f})")
    return d_lat < 2e-6 and d_hid < 5e-6


def run():
    args = make_args("highway", use_cuda=True, device="cuda")
    torch.manual_seed(3)
    pol = Behavior_policy(args, None)
    with torch.no_grad():
        pol.stack.flat.mul_(3.0)                 # larger weights than the default init: gates away from their linear range
    old = _lib.lib.iplan_behavior_get_impl()
    ok = True
    for B, mode in ((130, "contiguous"), (7, "strided"), (33, "strided_pad"), (512, "contiguous")):
        ok &= run_pair(pol, B, mode, seed=B)
    _lib.lib.iplan_behavior_set_impl(old)
    return ok


def timing(envs):
    args = make_args("highway", use_cuda=True, device="cuda")
    pol = Behavior_policy(args, None)
    A, N, o, L, W, E = args.n_agents, args.max_vehicle_num, args.obs_shape_single, args.latent_dim, args.max_history_len, args.encoder_rnn_dim
    win = torch.rand(A, envs, N, W * o, device="cuda") * 2 - 1
    hid = torch.zeros(A, envs, N, E, device="cuda")
    prev = torch.full((A, envs, N, L), 1.0 / L, device="cuda")
    new = torch.empty_like(prev)
    old = _lib.lib.iplan_behavior_get_impl()
    for impl in (1, 0):
        _lib.lib.iplan_behavior_set_impl(impl)
        for _ in range(3):
            pol.behavior_step(win, hid, prev, new)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            pol.behavior_step(win, hid, prev, new)
        e1.record()
        torch.cuda.synchronize()
        print(f"K1b impl {impl} ({'tcgen05' if impl == 0 else 'mma.sync'}) at {envs} envs: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per launch")
    _lib.lib.iplan_behavior_set_impl(old)
    if os.environ.get("IPLAN_BEH_DBG"):
        import ctypes as C
        clk = (C.c_longlong * 64)()
        _lib.check(_lib.lib.iplan_behavior_debug_clocks(clk), "debug_clocks")
        v = list(clk)
        t0 = v[0]
        print("K1b tcgen05 timeline of one CTA (cycles from kernel entry):")
        print("  tmem alloc %d | staged %d | synced %d | first products issued %d" % tuple(v[k] - t0 for k in (1, 2, 3, 4)))
        for w in range(10):
            a_, b_, c_, d_ = (v[8 + 4 * w + k] - t0 for k in range(4))
            print(f"  step {w}: top {a_}  input layer done +{b_ - a_}  products complete +{c_ - b_}  gates + operand stores +{d_ - c_}")
        print("  loop end %d | epilogue done %d | exit %d" % tuple(v[k] - t0 for k in (50, 51, 52)))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--time-envs", type=int, default=0)
    a = ap.parse_args()
    good = run()
    print("OK" if good else "MISMATCH")
    if a.time_envs:
        timing(a.time_envs)
    sys.exit(0 if good else 1)
