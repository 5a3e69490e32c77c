"""Compare the tcgen05 fc1 forward (iplan_learner_fc1_forward_tc5) with the mma.sync one
(iplan_learner_fc1_forward) on the same operands, and time both.  Run on a B200:

    timeout 120 python tools/check_fc1_tc5.py            # Highway shape + a ragged small case
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from iplan_b200 import _lib                          # noqa: E402
from iplan_b200.modules.flat import ParamStack       # noqa: E402


def run(A, rows, F, n_act=5, reps=5):
    torch.manual_seed(0)
    dev = "cuda"
    Fp = (F + 31) // 32 * 32
    actor = ParamStack("actor", A, (F, n_act), device=dev)
    critic = ParamStack("critic", A, (F,), device=dev)
    for st in (actor, critic):
        st.flat.copy_(torch.randn_like(st.flat) * 0.1)
    X = torch.zeros(A, rows, Fp, device=dev)
    X[..., :F] = torch.rand(A, rows, F, device=dev) * 2 - 1
    P, lib, st = _lib.ptr, _lib.lib, _lib.stream()
    z = lambda *s, **k: torch.zeros(*s, device=dev, **k)
    hz = lambda *s: torch.zeros(*s, device=dev, dtype=torch.float16)
    stat, Xh, Xl = z(A, rows, 2), hz(A, rows, Fp), hz(A, rows, Fp)
    Wh, Wl, ws, cc = hz(A, 128, Fp), hz(A, 128, Fp), z(A, 128), z(A, 128)
    _lib.check(lib.iplan_learner_row_stats(P(X), X.stride(0), Fp, F, rows, A, P(stat), st), "row_stats")
    _lib.check(lib.iplan_learner_x_split(P(X), X.numel(), P(Xh), P(Xl), st), "x_split")
    out = {}
    for name in ("iplan_learner_fc1_forward", "iplan_learner_fc1_forward_tc5"):
        fn = getattr(lib, name)
        Z = torch.full((A, rows, 128), float("nan"), device=dev)
        call = lambda: _lib.check(fn(P(actor.flat), actor.stride(), P(critic.flat), critic.stride(), P(Xh), P(Xl),
                                     Xh.stride(0), Fp, F, rows, A, P(stat), P(Wh), P(Wl), P(ws), P(cc), P(Z), st), name)
        call()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            call()
        e1.record()
        torch.cuda.synchronize()
        out[name] = (Z.clone(), e0.elapsed_time(e1) / reps)
    # fp64 reference of the folded product
    g = torch.stack([torch.cat([n.state_dict()["base.feature_norm.weight"] for n in (actor.nets[a],)]) for a in range(A)])
    za, zt = out["iplan_learner_fc1_forward"][0], out["iplan_learner_fc1_forward_tc5"][0]
    d = (za - zt).abs().max().item()
    ref = None
    xs = X[..., :F].double()
    mu, var = xs.mean(-1, keepdim=True), xs.var(-1, unbiased=False, keepdim=True)
    xn = (xs - mu) / torch.sqrt(var + 1e-5)
    zr = torch.empty(A, rows, 128, dtype=torch.float64, device=dev)
    for a in range(A):
        for t, stack in enumerate((actor, critic)):
            sd = stack.nets[a].state_dict()
            y = xn[a] * sd["base.feature_norm.weight"].double() + sd["base.feature_norm.bias"].double()
            zr[a, :, 64 * t:64 * t + 64] = y @ sd["base.mlp.fc1.0.weight"].double().T + sd["base.mlp.fc1.0.bias"].double()
    print(f"A={A} rows={rows} F={F}: mma {out['iplan_learner_fc1_forward'][1]:.3f} ms, tc5 {out['iplan_learner_fc1_forward_tc5'][1]:.3f} ms; "
          f"max|mma - tc5| = {d:.3e}; max|mma - fp64| = {(za.double() - zr).abs().max().item():.3e}; "
          f"max|tc5 - fp64| = {(zt.double() - zr).abs().max().item():.3e}; nan in tc5: {bool(torch.isnan(zt).any())}", flush=True)
    # ---- backward: G = dZ1^T X and the fc1.weight / feature_norm gradients -------------------------------
    dZ = torch.randn(A, rows, 128, device=dev) * 1e-3
    SM = torch.randn(A, 2, 128, device=dev) * 1e-3
    Dh, Dl, gs, G = hz(A, rows, 128), hz(A, rows, 128), z(2 * A), z(A, 128, Fp)
    res = {}
    for name in ("iplan_learner_fc1_backward", "iplan_learner_fc1_backward_tc5"):
        fn = getattr(lib, name)
        ga, gc = torch.zeros_like(actor.flat), torch.zeros_like(critic.flat)
        call = lambda: _lib.check(fn(P(actor.flat), actor.stride(), P(critic.flat), critic.stride(), P(ga), P(gc), P(Xh), P(Xl),
                                     Xh.stride(0), Fp, F, rows, A, P(dZ), P(Dh), P(Dl), P(gs), P(SM), P(G), st), name)
        call()
        torch.cuda.synchronize()
        g_first = (G.clone(), ga.clone(), gc.clone())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            call()
        e1.record()
        torch.cuda.synchronize()
        res[name] = (g_first, e0.elapsed_time(e1) / reps)
    Gref = torch.einsum("ark,arf->akf", dZ.double(), X.double())
    (Gm, gam, gcm), tm = res["iplan_learner_fc1_backward"]
    (Gt, gat, gct), tt = res["iplan_learner_fc1_backward_tc5"]
    scale = Gref.abs().max().item()
    db = max((Gm - Gt).abs().max().item(), (gam - gat).abs().max().item(), (gcm - gct).abs().max().item()) / scale
    print(f"   backward (all launches): mma {tm:.3f} ms, tc5 {tt:.3f} ms; max rel|mma - tc5| = {db:.3e}; "
          f"max rel|mma - fp64| = {(Gm.double() - Gref).abs().max().item() / scale:.3e}; "
          f"max rel|tc5 - fp64| = {(Gt.double() - Gref).abs().max().item() / scale:.3e}; nan: {bool(torch.isnan(Gt).any())}", flush=True)
    return max(d, db)


if __name__ == "__main__":
    ok = run(2, 300, 272) < 1e-4            # MPE width: ragged k (Fp = 288) and ragged rows
    ok &= run(1, 128, 2485) < 1e-4
    ok &= run(5, 46592, 2485, reps=3) < 1e-4   # bench shape: 512 envs x 91 steps
    print("OK" if ok else "MISMATCH")
