"""Time the tcgen05 recurrence kernel of K1 (impl 2) at the bench shape under the current IPLAN_GAT_DBG experiment bits
(results of the experiment variants are garbage: this is for timing only).   python tools/k1_variants.py [B]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from iplan_b200 import _lib                          # noqa: E402
from iplan_b200.modules.flat import ParamStack       # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
impl = int(os.environ.get("K1_IMPL", "2"))
A, N, o, L = 5, 55, 5, 8
torch.manual_seed(1)
dev = "cuda"
stack = ParamStack("gat", A, (o + L,), device=dev)
stack.flat.copy_(torch.randn_like(stack.flat) * 0.25)
hist = torch.rand(A, B, N, o, device=dev) * 2 - 1
beh = torch.softmax(torch.randn(A, B, N, L, device=dev), -1)
hprev = torch.rand(A, B, N, 32, device=dev) * 2 - 1
gum = -torch.log(torch.empty(A, B, N, N - 1, 2, device=dev).exponential_())
need = _lib.lib.iplan_gat_scratch_floats(B, A, N)
_lib.check(_lib.lib.iplan_gat_set_impl(impl), "set_impl")
scratch = torch.zeros(need, device=dev)
out = torch.zeros(A, B, N, 32, device=dev)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
for e in ev:
    e.record()
ts = []
for r in range(8):
    rc = _lib.lib.iplan_gat_step_ex(_lib.ptr(stack.flat), stack.stride(), _lib.view(hist), _lib.view(beh), _lib.view(hprev),
                                    _lib.view(out), _lib.ptr(gum), 1, 0, 0.01, None, _lib.ptr(scratch), need, B, A, N, o, L,
                                    ev[0].cuda_event, ev[1].cuda_event, ev[2].cuda_event, _lib.stream())
    _lib.check(rc, "gat_step_ex")
    torch.cuda.synchronize()
    ts.append(ev[0].elapsed_time(ev[1]) + (ev[1].elapsed_time(ev[2]) if impl == 0 else 0.0))
print(f"K1 impl {impl} dbg={os.environ.get('IPLAN_GAT_DBG', '0')} B={B}: min {min(ts[2:]):.4f} ms  median {sorted(ts[2:])[len(ts[2:]) // 2]:.4f} ms", flush=True)
