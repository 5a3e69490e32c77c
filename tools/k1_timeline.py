"""Phase timeline of one CTA of the fused K1 kernel (timing experiment; the stamped build is IPLAN_GAT_DBG=4):

    IPLAN_GAT_DBG=4 timeout 120 python tools/k1_timeline.py [B]

Prints clock64() deltas between the G5_STAMP points of csrc/gat_tc5.cu for CTA (3, 0), in SM cycles and as a share.
"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from iplan_b200 import _lib                          # noqa: E402
from iplan_b200.modules.flat import ParamStack       # noqa: E402

NAMES = ["start", "weights staged", "enc tile", "[P|Q] products", "Q table", "recurrence (this thread)", "recurrence (CTA)",
         "attn tiles staged", "qkv + gh products", "k / v tables", "scores + hard gate", "soft-max", "aggregation", "x tile",
         "gi product", "gates + store"]

if __name__ == "__main__":
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    A, N, o, L = 5, 55, 5, 8
    dev = "cuda"
    stack = ParamStack("gat", A, (o + L,), device=dev)
    stack.flat.copy_(torch.randn_like(stack.flat) * 0.25)
    hist = torch.rand(A, B, N, o, device=dev) * 2 - 1
    beh = torch.softmax(torch.randn(A, B, N, L, device=dev), -1)
    hprev = torch.rand(A, B, N, 32, device=dev) * 2 - 1
    need = _lib.lib.iplan_gat_scratch_floats(B, A, N)
    scratch = torch.zeros(need, device=dev)
    out = torch.zeros(A, B, N, 32, device=dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    for e in ev:
        e.record()
    _lib.check(_lib.lib.iplan_gat_set_impl(0), "set_impl")
    for r in range(4):
        rc = _lib.lib.iplan_gat_step_ex(_lib.ptr(stack.flat), stack.stride(), _lib.view(hist), _lib.view(beh), _lib.view(hprev),
                                        _lib.view(out), None, 1, r, 0.01, None, _lib.ptr(scratch), need, B, A, N, o, L,
                                        ev[0].cuda_event, ev[1].cuda_event, ev[2].cuda_event, _lib.stream())
        _lib.check(rc, "gat_step_ex")
        torch.cuda.synchronize()
    print(f"kernel {ev[0].elapsed_time(ev[2]):.4f} ms (stamped build)")
    clk = (ctypes.c_longlong * 32)()
    _lib.check(_lib.lib.iplan_gat_debug_clocks(clk), "debug_clocks")
    c = list(clk)[:16]
    tot = c[15] - c[0]
    for k in range(1, 16):
        print(f"  {NAMES[k]:28s} {c[k] - c[k - 1]:8d} cycles  {100.0 * (c[k] - c[k - 1]) / tot:5.1f} %")
    print(f"  total {tot} cycles; recurrence per step {(c[5] - c[4]) / (N - 1):.0f} cycles")
