"""IPLAN_GAT_DBG=64: per-step event timeline (clock64, relative) of warps 0 / 4 / 8 / 12 of one CTA of the tcgen05 K1 kernel, steps
20..27, and the CTA's phase stamps.   [K1_IMPL=0|2] python tools/k1_trace.py      (0 = the fused kernel, 2 = recurrence only)"""
import ctypes
import os
import sys

os.environ.setdefault("IPLAN_GAT_DBG", "64")
os.environ.setdefault("K1_IMPL", "2")
sys.argv = [sys.argv[0], "512"]
exec(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "k1_variants.py")).read())
buf = (ctypes.c_longlong * 256)()
_lib.check(_lib.lib.iplan_gat_debug_trace(buf), "trace")
t0 = min(v for v in buf if v > 0)
names = ["top", "mbar", "ld0", "sts0", "sts1", "arrive"]
for w in range(4):
    print(f"warp {4 * w} (tile {w & 1}, half {w >> 1})")
    for st in range(8):
        ev = [buf[(w * 8 + st) * 8 + e] for e in range(6)]
        print(f"  step {20 + st}: " + "  ".join(f"{n}={ev[k] - t0:7d}" for k, n in enumerate(names)))
clk = (ctypes.c_longlong * 32)()
_lib.check(_lib.lib.iplan_gat_debug_clocks(clk), "clocks")
st = [(k, clk[k]) for k in range(16) if clk[k] > 0]
if st:
    print("phase stamps (cycles since the CTA's first stamp; 0 start, 1 operand tiles staged, 2 pipeline primed, 3 recurrence done, "
          "4.. attention phase syncs, 14 GRUCell products done, 15 end):")
    print("  " + "  ".join(f"[{k}]={v - st[0][1]}" for k, v in st))
