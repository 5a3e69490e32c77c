"""TEST / BENCH INFRASTRUCTURE — the timed CPU baseline.

Drives the oracle port (oracle/iplan_oracle.py) in the reference's own call order and op
structure on the host cores:

  per timestep (runners/ippo_parallel_runner.py:166-266):
      DcntrlMAC.select_actions_ippo -> Prediction_policy.GAT_latent_update (per-agent Python
      loop, per-ego cat/stack loops, bidirectional nn.GRU) -> Behavior_policy.latent_update,
      with numpy <-> torch conversions at each module boundary as the reference does
  per update (learners/ippo_learner.py:227-317): per agent, 15 epochs of autograd + Adam.

/root/reference is not present on the GPU box, so this is a "port" baseline
(cpu_baseline.kind = "port"); tests/test_oracle_golden.py pins the port to the reference's
own outputs.  Used only by bench.py.
"""
import os
import time

import numpy as np
import torch

from . import iplan_oracle as O


def _torch_default_init(shape, fan):
    b = 1.0 / np.sqrt(fan)
    return (torch.rand(*shape) * 2 - 1) * b


def random_params(args, seed=0):
    """Weights with the reference modules' shapes and init scales (values are irrelevant
    for timing; shapes and sparsity of the inputs are what matter)."""
    g = torch.Generator().manual_seed(seed)
    A, N, o, L, H, R = args.n_agents, args.max_vehicle_num, args.obs_shape_single, args.latent_dim, 32, 64
    F = N * (o + H + L) + args.n_actions + A

    def u(*shape, fan):
        return (torch.rand(*shape, generator=g) * 2 - 1) / np.sqrt(fan)

    gat, beh, act, cri = [], [], [], []
    for _ in range(A):
        p = {"encoding.weight": u(H, o + L, fan=o + L), "encoding.bias": u(H, fan=o + L)}
        for sfx in ("", "_reverse"):
            p["hard_bi_GRU.weight_ih_l0" + sfx] = u(3 * H, 2 * H, fan=H)
            p["hard_bi_GRU.weight_hh_l0" + sfx] = u(3 * H, H, fan=H)
            p["hard_bi_GRU.bias_ih_l0" + sfx] = u(3 * H, fan=H)
            p["hard_bi_GRU.bias_hh_l0" + sfx] = u(3 * H, fan=H)
        p.update({"hard_encoding.weight": u(2, 2 * H, fan=2 * H), "hard_encoding.bias": u(2, fan=2 * H),
                  "q.weight": u(H, H, fan=H), "k.weight": u(H, H, fan=H), "v.weight": u(H, H, fan=H), "v.bias": u(H, fan=H),
                  "rnn.weight_ih": u(3 * H, H, fan=H), "rnn.weight_hh": u(3 * H, H, fan=H),
                  "rnn.bias_ih": u(3 * H, fan=H), "rnn.bias_hh": u(3 * H, fan=H)})
        gat.append(p)
        beh.append({"linear.weight": u(H, o, fan=o), "linear.bias": u(H, fan=o),
                    "rnn.weight_ih_l0": u(3 * H, H, fan=H), "rnn.weight_hh_l0": u(3 * H, H, fan=H),
                    "rnn.bias_ih_l0": u(3 * H, fan=H), "rnn.bias_hh_l0": u(3 * H, fan=H),
                    "out.weight": u(L, H, fan=H), "out.bias": u(L, fan=H)})

        def trunk():
            t = {"base.feature_norm.weight": torch.ones(F), "base.feature_norm.bias": torch.zeros(F),
                 "base.mlp.fc1.0.weight": u(R, F, fan=F), "base.mlp.fc1.0.bias": torch.zeros(R),
                 "base.mlp.fc1.2.weight": torch.ones(R), "base.mlp.fc1.2.bias": torch.zeros(R),
                 "base.mlp.fc_h.0.weight": u(R, R, fan=R), "base.mlp.fc_h.0.bias": torch.zeros(R),
                 "base.mlp.fc_h.2.weight": torch.ones(R), "base.mlp.fc_h.2.bias": torch.zeros(R),
                 "base.mlp.fc2.0.0.weight": u(R, R, fan=R), "base.mlp.fc2.0.0.bias": torch.zeros(R),
                 "base.mlp.fc2.0.2.weight": torch.ones(R), "base.mlp.fc2.0.2.bias": torch.zeros(R),
                 "rnn.rnn.weight_ih_l0": u(3 * R, R, fan=R), "rnn.rnn.weight_hh_l0": u(3 * R, R, fan=R),
                 "rnn.rnn.bias_ih_l0": torch.zeros(3 * R), "rnn.rnn.bias_hh_l0": torch.zeros(3 * R),
                 "rnn.norm.weight": torch.ones(R), "rnn.norm.bias": torch.zeros(R)}
            return t
        a = trunk()
        a.update({"act.action_out.linear.weight": u(args.n_actions, R, fan=R) * 0.1,
                  "act.action_out.linear.bias": torch.zeros(args.n_actions)})
        c = trunk()
        c.update({"v_out.weight": u(1, R, fan=R), "v_out.bias": torch.zeros(1)})
        act.append(a)
        cri.append(c)
    return dict(gat=gat, beh=beh, actors=act, critics=cri, F=F)


def synth_step(rng, B, A, N, o, k):
    h = rng.uniform(-1, 1, size=(B, A, N, o))
    h[..., 0] = 1.0
    h[:, :, k:] = 0.0
    return h


def time_rollout_steps(args, params, B, n_steps, warmup=1, seed=0):
    """Seconds per rollout timestep at B envs (mean over n_steps after `warmup`)."""
    rng = np.random.default_rng(seed)
    A, N, o, L, D, E, W, R = (args.n_agents, args.max_vehicle_num, args.obs_shape_single, args.latent_dim,
                              args.attention_dim, args.encoder_rnn_dim, args.max_history_len, args.rnn_hidden_dim)
    att = np.zeros((B, A, N, D), dtype=np.float32)
    beh = np.zeros((B, A, N, L), dtype=np.float32)
    enc = np.zeros((B, 1, A, N, E), dtype=np.float32)
    rnn_a = torch.zeros(B, A, R)
    rnn_c = torch.zeros(B, A, R)
    last = torch.zeros(B, A, args.n_actions)
    avail = torch.ones(B, A, args.n_actions)
    times = []
    for it in range(warmup + n_steps):
        single = synth_step(rng, B, A, N, o, min(N, getattr(args, "n_obs_vehicles", N) + it))
        window = rng.uniform(-1, 1, size=(B, A, N, W, o))
        t0 = time.perf_counter()
        with torch.no_grad():
            x = O.build_inputs_step(torch.as_tensor(single, dtype=torch.float32), torch.as_tensor(att),
                                    torch.as_tensor(beh), last, A)
            r = O.select_actions(params["actors"], params["critics"], x, avail, rnn_a, rnn_c,
                                 uniforms=torch.rand(B, A))
            actions = r["actions"].numpy()
            rnn_a, rnn_c = r["rnn_a"], r["rnn_c"]
            gum = torch.stack([O.draw_gumbel(B * N * (N - 1)).view(B, N, N - 1, 2) for _ in range(A)])
            att = O.gat_latent_update(params["gat"], single, att, beh, gum, loops=True).numpy()
            beh_t, enc_t = O.behavior_latent_update(params["beh"], window, enc, beh, args.soft_update_coef)
            beh, enc = beh_t.numpy(), enc_t.numpy()
            last = torch.nn.functional.one_hot(torch.as_tensor(actions), args.n_actions).float()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return float(np.mean(times))


def time_train(args, params, Bf, seed=0):
    """Seconds for one IPPOLearner.train-equivalent over Bf episodes of T steps (all agents)."""
    import copy
    from types import SimpleNamespace
    rng = np.random.default_rng(seed)
    A, N, o, L, D, R, T = (args.n_agents, args.max_vehicle_num, args.obs_shape_single, args.latent_dim,
                           args.attention_dim, args.rnn_hidden_dim, args.episode_limit)
    a2 = SimpleNamespace(**vars(args))
    a2.buffer_size, a2.batch_size = Bf, Bf - 1
    t0 = time.perf_counter()
    for a in range(A):
        hist = rng.uniform(-1, 1, size=(Bf, T + 1, N, o)).astype(np.float32)
        hist[:, :, 30:] = 0
        actions = torch.as_tensor(rng.integers(0, args.n_actions, size=(Bf, T + 1, 1)))
        batch = dict(history=torch.as_tensor(hist),
                     attention_latent=torch.as_tensor(rng.uniform(-1, 1, size=(Bf, T + 1, N, D)).astype(np.float32)),
                     behavior_latent=torch.as_tensor(rng.uniform(0, 1, size=(Bf, T + 1, N, L)).astype(np.float32)),
                     actions=actions, actions_onehot=torch.nn.functional.one_hot(actions.squeeze(-1), args.n_actions).float(),
                     available_actions=torch.ones(Bf, T + 1, args.n_actions),
                     reward=torch.as_tensor(rng.normal(size=(Bf, T + 1, 1)).astype(np.float32)),
                     terminated_masks=torch.ones(Bf, T + 1, 1),
                     rnn_states_actor=torch.zeros(Bf, T + 1, R), rnn_states_critic=torch.zeros(Bf, T + 1, R))
        ap = {k: v.clone() for k, v in params["actors"][a].items()}
        cp = {k: v.clone() for k, v in params["critics"][a].items()}
        O.train_agent(ap, cp, batch, a, a2)
    return time.perf_counter() - t0


def usable_cpus(cap=32):
    """Host threads torch may use: the scheduler affinity, clipped by the cgroup CPU quota (a
    container that shows 128 cores but owns 8 of them crawls when 128 OpenMP threads spin)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = min(n, max(1, int(int(quota) / int(period))))
    except Exception:
        pass
    return max(1, min(n, cap))


def _log(msg):
    import sys
    print(f"[cpu_baseline {time.strftime('%H:%M:%S')}] {msg}", file=sys.stderr, flush=True)


def measure(args, B=512, rollout_steps=2, train_eps=32, threads=None, budget_s=25.0):
    """env-steps/s of the CPU port for the B-env workload, on a BOUNDED sample: a probe timestep
    at 16 envs sizes the sample so that rollout timing (1 warm-up + `rollout_steps` timesteps at
    B_s <= B envs) and one update at `train_eps` episodes fit ~`budget_s` seconds.  Rollout cost
    per env is taken from the B_s-env timing (it improves slowly with B: the reference's per-op
    overhead amortises), the update is scaled linearly in rows (GEMMs and row-wise ops over Bf*T)."""
    threads = threads or usable_cpus()
    torch.set_num_threads(threads)
    params = random_params(args)
    T = args.episode_limit
    t0 = time.perf_counter()
    probe = time_rollout_steps(args, params, 16, 1, warmup=0)
    _log(f"{threads} threads; probe timestep at 16 envs: {probe:.2f} s")
    Bs = B
    while Bs > 16 and probe * (Bs / 16) * (1 + rollout_steps) > 0.6 * budget_s:
        Bs //= 2
    t_step = time_rollout_steps(args, params, Bs, rollout_steps, warmup=1)
    _log(f"rollout timestep at {Bs} envs: {t_step:.2f} s")
    eps = train_eps
    while eps > 4 and (time.perf_counter() - t0) + 0.5 * eps > budget_s * 1.5:
        eps //= 2
    t_train_small = time_train(args, params, eps)
    _log(f"update at {eps} episodes: {t_train_small:.2f} s")
    t_train = t_train_small * (B / eps)
    value = B * T / (T * t_step * (B / Bs) + t_train)
    sample = (f"rollout: 1 warm-up + {rollout_steps} timed timesteps at {Bs} envs (scaled x{B / Bs:g} to {B} envs, x{T} "
              f"timesteps); update: one train() at Bf={eps} episodes (T={T}, {args.ppo_epoch} epochs, {args.n_agents} agents) "
              f"scaled x{B / eps:g} in rows; {threads} torch threads")
    return dict(value=value, t_step=t_step * (B / Bs), t_train=t_train, cores=threads, sample=sample)
