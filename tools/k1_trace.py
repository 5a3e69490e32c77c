"""IPLAN_GAT_DBG=64: print the per-step event timeline (clock64, relative) of warps 0 / 4 / 8 / 12 of one CTA of the
tcgen05 recurrence kernel, steps 20..27.   IPLAN_GAT_DBG=64 python tools/k1_trace.py"""
import ctypes
import os
import sys

os.environ.setdefault("IPLAN_GAT_DBG", "64")
os.environ["K1_IMPL"] = "2"
sys.argv = [sys.argv[0], "512"]
exec(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "k1_variants.py")).read())
buf = (ctypes.c_longlong * 256)()
_lib.check(_lib.lib.iplan_gat_debug_trace(buf), "trace")
t0 = min(v for v in buf if v > 0)
names = ["top", "mbar", "ld0", "sts0", "sts1", "bar", "issued"]
for w in range(4):
    print(f"warp {4 * w} (tile {w & 1}, half {w >> 1})")
    for st in range(8):
        ev = [buf[(w * 8 + st) * 8 + e] for e in range(7)]
        print(f"  step {20 + st}: " + "  ".join(f"{n}={ev[k] - t0:7d}" for k, n in enumerate(names)))
