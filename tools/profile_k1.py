"""Launch K1 (iplan_gat_step) a few times on the bench shape and nothing else, so that an
``ncu --set full -k regex:gat_ -s 4 -c 2 python tools/profile_k1.py`` capture costs seconds.
Also prints CUDA-event timings of the pair of launches (not valid under ncu).

Inputs follow the rollout's statistics: bounded histories, softmax behaviour latents, tanh-range
hidden state (the kernel's duration is data independent)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from iplan_b200.config import make_args            # noqa: E402
from iplan_b200.nova.prediction_policy import Prediction_policy   # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    args = make_args("highway", batch_size_run=B, use_cuda=True, device="cuda")
    A, N, o, L = args.n_agents, args.max_vehicle_num, args.obs_shape_single, args.latent_dim
    torch.manual_seed(0)
    pol = Prediction_policy(args)
    hist = torch.rand(A, B, N, o, device="cuda") * 2 - 1
    beh = torch.softmax(torch.randn(A, B, N, L, device="cuda"), -1)
    h = torch.tanh(torch.randn(A, B, N, 32, device="cuda"))
    out = torch.empty_like(h)
    for _ in range(3):
        pol.gat_step(hist, beh, h, out)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        pol.gat_step(hist, beh, out if i % 2 == 0 else h, h if i % 2 == 0 else out)
        ev[i + 1].record()
    torch.cuda.synchronize()
    ms = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))
    print(f"K1 pair of launches, B={B}: median {ms[len(ms) // 2]:.3f} ms, min {ms[0]:.3f} ms over {reps}")


if __name__ == "__main__":
    main()
