"""Compare the tcgen05 kernels of K1 (impl 0: one fused launch; impl 2: tcgen05 recurrence + attention kernel) with the
mma.sync path (impl 1) on the same inputs -- the dl scratch both hand to the attention kernel and the final output -- and time the
recurrence kernel of each with the CUDA events of iplan_gat_step_ex.  Run on a B200:

    timeout 300 python tools/check_gat_tc5.py            # B = 3, 6 (ragged CTA pairs), 64, 512 at the Highway shape; MPE shape
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from iplan_b200 import _lib                          # noqa: E402
from iplan_b200.modules.flat import ParamStack       # noqa: E402


def run(B, A, N, o=5, L=8, reps=5, label=""):
    torch.manual_seed(B * 131 + N)
    dev = "cuda"
    stack = ParamStack("gat", A, (o + L,), device=dev)
    stack.flat.copy_(torch.randn_like(stack.flat) * 0.25)
    hist = torch.rand(A, B, N, o, device=dev) * 2 - 1
    beh = torch.softmax(torch.randn(A, B, N, L, device=dev), -1)
    hprev = torch.rand(A, B, N, 32, device=dev) * 2 - 1
    gum = -torch.log(torch.empty(A, B, N, N - 1, 2, device=dev).exponential_())
    need = _lib.lib.iplan_gat_scratch_floats(B, A, N)
    res = {}
    for impl in (1, 2, 0):
        _lib.check(_lib.lib.iplan_gat_set_impl(impl), "set_impl")
        scratch = torch.zeros(need, device=dev)
        out = torch.zeros(A, B, N, 32, device=dev)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        for e in ev:
            e.record()                                  # create the handles
        times = []
        for r in range(reps):
            rc = _lib.lib.iplan_gat_step_ex(_lib.ptr(stack.flat), stack.stride(), _lib.view(hist), _lib.view(beh),
                                            _lib.view(hprev), _lib.view(out), _lib.ptr(gum), 1, 0, 0.01, None,
                                            _lib.ptr(scratch), need, B, A, N, o, L,
                                            ev[0].cuda_event, ev[1].cuda_event, ev[2].cuda_event, _lib.stream())
            _lib.check(rc, "gat_step_ex")
            torch.cuda.synchronize()
            times.append((ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])))
        res[impl] = (scratch.view(A, B, 2, N - 1, 64)[..., :N].clone(), out.clone(), times)
    dl1, out1, t1 = res[1]
    dl2, out2, t2 = res[2]
    _, out0, t0 = res[0]                        # fused: no dl scratch
    ddl = float((dl1 - dl2).abs().max())
    dout = max(float((out1 - out2).abs().max()), float((out1 - out0).abs().max()))
    nan = bool(torch.isnan(dl2).any() or torch.isnan(out2).any() or torch.isnan(out0).any())
    best = lambda ts, k: min(t[k] for t in ts[1:] or ts)
    print(f"[{label} B={B} A={A} N={N}] max|dl tc5 - dl mma| = {ddl:.3e} (|dl| max {float(dl1.abs().max()):.2f})  "
          f"max|out - out mma|: tc5+attend {float((out1 - out2).abs().max()):.3e}, fused {float((out1 - out0).abs().max()):.3e}  nan={nan}\n"
          f"      ms: mma recur {best(t1, 0):.4f} + attend {best(t1, 1):.4f} | tc5 recur {best(t2, 0):.4f} + attend {best(t2, 1):.4f} | "
          f"fused {best(t0, 0) + best(t0, 1):.4f}", flush=True)
    return ddl, dout, nan


if __name__ == "__main__":
    assert torch.cuda.is_available()
    bad = 0
    for (B, A, N, o, L, lab) in ((3, 5, 55, 5, 8, "highway"), (6, 5, 55, 5, 8, "highway"), (4, 3, 6, 4, 8, "mpe"),
                                 (64, 5, 55, 5, 8, "highway"), (512, 5, 55, 5, 8, "highway")):
        ddl, dout, nan = run(B, A, N, o, L, label=lab)
        bad += int(nan or ddl > 1e-4 or dout > 1e-4)
    print("CHECK", "FAILED" if bad else "OK")
    sys.exit(1 if bad else 0)
