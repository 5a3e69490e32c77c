"""TEST / BENCH INFRASTRUCTURE — drives the REAL reference (wuxiyang1996/iPLAN), staged UNMODIFIED under oracle/_ref/ by
oracle/make_ref.py, on the host cores.

Nothing here restates the reference's arithmetic: `DcntrlMAC`, `IPPOLearner`, `Prediction_policy`, `Behavior_policy`,
`EpisodeBatch`, `ParallelRunner` and `observersation_state_history_wrapper` are the reference's own classes imported from
oracle/_ref.  This module only supplies what the reference gets from files that are out of scope: the YAML
merge of main.py:59-100, the `args` namespace of run_ippo.py:39,136-147, the scheme of run_ippo.py:160-184, a logger stub
and a seeded stand-in for the simulator (SubprocVecEnv over Heterogeneous_Highway_Env).

Used by: bench.py (`--impl reference`, and the `cpu_baseline` leg; kind "reference") and tests/test_gpu_dropin.py.
"""
import json
import os
import sys
import time
from types import SimpleNamespace

import numpy as np
import torch
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def available():
    return os.path.exists(os.path.join(REF_DIR, "MANIFEST.json"))


def activate():
    """Make the staged reference importable (its modules use absolute top-level names: controllers, learners, ...)."""
    if not available():
        raise RuntimeError("oracle/_ref is not staged: run `python oracle/make_ref.py` in the build container")
    sys.dont_write_bytecode = True
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)


def manifest():
    with open(os.path.join(REF_DIR, "MANIFEST.json")) as f:
        return json.load(f)


# ---- what main.py / run_ippo.py do before the hot path starts ------------------------------------------------
def load_config(env="highway"):
    """main.py:59-69, :83-100: default.yaml, then the env file, then the algorithm file; keys already present win."""
    def load(*p):
        with open(os.path.join(REF_DIR, "config", *p)) as f:
            return yaml.safe_load(f)
    cfg = load("default.yaml")
    cfg["env"] = env                                        # the key the user edits to pick the env file (main.py:85-92)
    for layer in (load("envs", "highway.yaml" if env == "highway" else "simple_spread_Hetero.yaml"), load("algs", "ippo.yaml")):
        for k, v in layer.items():
            if k not in cfg:
                cfg[k] = v
    return cfg


def ref_args(env="highway", **over):
    """The namespace run_ippo.run / run_sequential hold once the runner has reported its env_info (run_ippo.py:39,136-147)."""
    cfg = load_config(env)
    cfg.update(over)
    a = SimpleNamespace(**cfg)
    a.use_cuda, a.device = False, "cpu"
    if a.env == "highway":
        a.max_vehicle_num = a.n_other_vehicles + a.n_agents
        a.state_shape = a.obs_shape_single * a.max_vehicle_num
        a.obs_shape = a.obs_shape_single * a.n_obs_vehicles
    else:
        a.n_agents = a.num_agents
        a.max_vehicle_num = a.num_landmarks + a.n_agents + a.num_random_agents
        a.episode_limit = a.episode_length
        a.state_shape = a.obs_shape_single * a.max_vehicle_num
        a.obs_shape = a.obs_shape_single * a.max_vehicle_num
    for k, v in over.items():
        setattr(a, k, v)
    return a


def make_scheme(args):
    """run_ippo.py:160-184."""
    activate()
    from components.transforms import OneHot
    scheme = {
        "state": {"vshape": args.state_shape},
        "obs": {"vshape": args.obs_shape, "group": "agents"},
        "actions": {"vshape": (1,), "group": "agents", "dtype": torch.long},
        "rnn_states_actors": {"vshape": (args.rnn_hidden_dim,), "group": "agents"},
        "rnn_states_critics": {"vshape": (args.rnn_hidden_dim,), "group": "agents"},
        "history": {"vshape": (args.max_vehicle_num, args.obs_shape_single,), "group": "agents"},
        "behavior_latent": {"vshape": (args.max_vehicle_num, args.latent_dim,), "group": "agents"},
        "attention_latent": {"vshape": (args.max_vehicle_num, args.attention_dim,), "group": "agents"},
        "avail_actions": {"vshape": (args.n_actions,), "group": "agents", "dtype": torch.int},
        "reward": {"vshape": (1,), "group": "agents"},
        "speed": {"vshape": (1,), "group": "agents"},
        "terminated": {"vshape": (1,), "group": "agents", "dtype": torch.uint8},
    }
    return scheme, {"agents": args.n_agents}, {"actions": ("actions_onehot", [OneHot(out_dim=args.n_actions)])}


class NullLogger:
    def __init__(self):
        self.stats = {}

    def log_stat(self, key, value, t):
        self.stats[key] = float(value)


class SyntheticHostEnv:
    """Seeded numpy stand-in for SubprocVecEnv(Heterogeneous_Highway_Env) with the interface ParallelRunner uses
    (runners/ippo_parallel_runner.py:96, :185): `reset() -> (state, obs)`, `step(actions) -> (state, obs, reward, win_tags,
    terminated_agent, env_info)`, `close()`.  Independent of the actions (the metric is defined on synthetic observations,
    SURVEY §8d).  obs [B, A, n_obs, 1 + o]: column 0 = vehicle id (ego first, observation_wrapper.py:73), then presence = 1
    and U(-1,1) features; the set of vehicles an agent has met grows by one every third step, so slots fill as
    K_t = min(N, n_obs + t // 3).  terminated latches per agent with a per-step hazard."""

    def __init__(self, args, n_envs, hazard=0.01, seed=112358):
        self.args, self.B, self.hazard = args, n_envs, hazard
        self.A, self.N, self.o, self.M = args.n_agents, args.max_vehicle_num, args.obs_shape_single, args.n_obs_vehicles
        self.rng = np.random.default_rng(seed)
        self.t = 0

    def _obs(self):
        B, A, M, o, N = self.B, self.A, self.M, self.o, self.N
        obs = np.zeros((B, A, M, o + 1))
        obs[..., 1] = 1.0
        obs[..., 2:] = self.rng.uniform(-1, 1, size=(B, A, M, o - 1))
        obs[:, :, 0, 0] = 1 + np.arange(A)[None, :]                      # ego ids 1..A
        first = self.t // 3                                              # window of the others' ids slides by one every 3 steps
        others = 100 + (first + np.arange(M - 1)) % (N - 1)
        obs[:, :, 1:, 0] = others[None, None, :]
        state = self.rng.uniform(-1, 1, size=(B, 1, N * (o + 1)))
        return state, obs

    def reset(self):
        self.t = 0
        self.dead = np.zeros((self.B, self.A), dtype=bool)
        return self._obs()

    def step(self, actions):
        self.t += 1
        state, obs = self._obs()
        reward = self.rng.normal(size=(self.B, self.A))
        self.dead |= self.rng.uniform(size=(self.B, self.A)) < self.hazard
        win = np.zeros((self.B, self.A))
        info = [{"speed": np.full(self.A, 20.0)} for _ in range(self.B)]
        return state, obs, reward, win, self.dead.copy(), info

    def close(self):
        pass


def build_reference(args, seed=0):
    """The objects run_sequential constructs (run_ippo.py:188-222), all from the staged reference."""
    activate()
    from components.episode_buffer import EpisodeBatch
    from controllers.dcntrl_controller import DcntrlMAC
    from learners.ippo_learner import IPPOLearner
    from nova.prediction_policy import Prediction_policy
    from nova.stable_behavior_policy import Behavior_policy
    torch.manual_seed(seed)
    logger = NullLogger()
    scheme, groups, preprocess = make_scheme(args)
    proto = EpisodeBatch(scheme, groups, 1, 2, preprocess=preprocess, device="cpu")
    mac = DcntrlMAC(proto.scheme, groups, args)
    learner = IPPOLearner(mac, proto.scheme, logger, args)
    beh = Behavior_policy(args, logger)
    pred = Prediction_policy(args, logger)
    return SimpleNamespace(args=args, logger=logger, scheme=scheme, groups=groups, preprocess=preprocess,
                           mac=mac, learner=learner, behavior=beh, prediction=pred)


def build_runner(sysm, env):
    activate()
    from runners.ippo_parallel_runner import ParallelRunner
    runner = ParallelRunner(sysm.args, env, sysm.logger)
    runner.setup(sysm.scheme, sysm.groups, sysm.preprocess, sysm.mac, sysm.behavior, sysm.prediction)
    return runner


# ---- timing ------------------------------------------------------------------------------------------------
def usable_cpus(cap=32):
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = min(n, max(1, int(int(quota) / int(period))))
    except Exception:
        pass
    return max(1, min(n, cap))


def _log(msg):
    print(f"[ref_driver {time.strftime('%H:%M:%S')}] {msg}", file=sys.stderr, flush=True)


class _Timers:
    """Wall time spent inside the reference's hot-path methods while its own runner body drives them."""

    def __init__(self):
        self.calls = []            # (name, seconds) in call order

    def wrap(self, obj, name, tag=None):
        fn = getattr(obj, name)
        tag = tag or name

        def timed(*a, **k):
            t0 = time.perf_counter()
            try:
                return fn(*a, **k)
            finally:
                self.calls.append((tag, time.perf_counter() - t0))
        setattr(obj, name, timed)


def time_rollout_timestep(B, timesteps=1, seed=0, hazard=0.01, warm=False):
    """Seconds of hot-path work per rollout timestep at B envs, measured inside the reference's OWN ParallelRunner.run
    (runners/ippo_parallel_runner.py:105-281) driving its own MAC / GAT / behaviour modules and EpisodeBatch on the
    synthetic env: select_actions_ippo + GAT_latent_update + latent_update + EpisodeBatch.update of one loop iteration.
    The runner's Python observation wrapper and the env are out of scope and not counted.
    `timesteps` loop iterations are timed (after one untimed iteration if `warm`)."""
    activate()
    from components.episode_buffer import EpisodeBatch
    args = ref_args("highway", batch_size_run=B, episode_limit=timesteps + (1 if warm else 0), buffer_size=B, batch_size=B - 1)
    sysm = build_reference(args, seed)
    runner = build_runner(sysm, SyntheticHostEnv(args, B, hazard, seed=112358 + seed))
    tm = _Timers()
    tm.wrap(sysm.mac, "select_actions_ippo")
    tm.wrap(sysm.prediction, "GAT_latent_update")
    tm.wrap(sysm.behavior, "latent_update")
    orig_update = EpisodeBatch.update

    def timed_update(self, *a, **k):
        t0 = time.perf_counter()
        try:
            return orig_update(self, *a, **k)
        finally:
            tm.calls.append(("batch_update", time.perf_counter() - t0))
    EpisodeBatch.update = timed_update
    try:
        runner.run(test_mode=False)
    finally:
        EpisodeBatch.update = orig_update
    # split the call log into loop iterations at each select_actions_ippo; iteration 0 is the warm-up
    iters, cur = [], None
    for name, dt in tm.calls:
        if name == "select_actions_ippo":
            cur = {}
            iters.append(cur)
        if cur is not None:
            cur[name] = cur.get(name, 0.0) + dt
    timed = iters[1:] if warm and len(iters) > 1 else iters
    per = {k: float(np.mean([it.get(k, 0.0) for it in timed])) for k in ("select_actions_ippo", "GAT_latent_update", "latent_update", "batch_update")}
    return sum(per.values()), per


def time_update(Bf, seed=0):
    """Seconds of the reference's IPPOLearner.insert_episode_batch + train (learners/ippo_learner.py:96-126, :227-317) over Bf
    full-length episodes (T = 90, 15 epochs, all agents)."""
    activate()
    from components.episode_buffer import EpisodeBatch
    args = ref_args("highway", batch_size_run=Bf, buffer_size=Bf, batch_size=Bf - 1)
    sysm = build_reference(args, seed)
    A, N, o, L, D, R, T = (args.n_agents, args.max_vehicle_num, args.obs_shape_single, args.latent_dim, args.attention_dim,
                           args.rnn_hidden_dim, args.episode_limit)
    rng = np.random.default_rng(seed)
    batch = EpisodeBatch(sysm.scheme, sysm.groups, Bf, T + 1, preprocess=sysm.preprocess, device="cpu")
    hist = rng.uniform(-1, 1, size=(Bf, T + 1, A, N, o)).astype(np.float32)
    hist[..., 0] = 1.0
    for t in range(T + 1):
        hist[:, t, :, min(N, args.n_obs_vehicles + t // 3):] = 0.0
    term = (np.cumsum(rng.uniform(size=(Bf, T + 1, A, 1)) < 0.01, axis=1) > 0).astype(np.uint8)
    batch.update(dict(history=hist,
                      attention_latent=rng.uniform(-1, 1, size=(Bf, T + 1, A, N, D)).astype(np.float32),
                      behavior_latent=rng.dirichlet(np.ones(L), size=(Bf, T + 1, A, N)).astype(np.float32),
                      rnn_states_actors=rng.uniform(-1, 1, size=(Bf, T + 1, A, R)).astype(np.float32),
                      rnn_states_critics=rng.uniform(-1, 1, size=(Bf, T + 1, A, R)).astype(np.float32),
                      actions=rng.integers(0, args.n_actions, size=(Bf, T + 1, A, 1)),
                      avail_actions=np.ones((Bf, T + 1, A, args.n_actions), dtype=np.int64),
                      reward=rng.normal(size=(Bf, T + 1, A, 1)).astype(np.float32), terminated=term),
                 bs=slice(None), ts=slice(None))
    t0 = time.perf_counter()
    sysm.learner.insert_episode_batch(batch)
    sysm.learner.train(t_env=0)
    dt = time.perf_counter() - t0
    assert len(sysm.logger.stats) >= 6, "the reference's train() returned without updating (buffer not full?)"
    return dt


ROLLOUT_ENVS = 512       # the fixed sample: ONE full rollout timestep at the benchmark's 512 envs ...
UPDATE_EPISODES = 64     # ... and one train() over 64 full-length episodes (rows scale x8 to the 512-episode update)


def measure(B=512, T=90, threads=None, seed=0, rollout_envs=ROLLOUT_ENVS, update_episodes=UPDATE_EPISODES):
    """env-steps/s of the reference's own CPU path for the B-env workload on a FIXED sample (no probe-dependent sizing):
    rollout = T x (one timed timestep at `rollout_envs` envs, scaled linearly in envs if rollout_envs != B);
    update  = train() at Bf = `update_episodes`, scaled linearly in rows (B / update_episodes): every op of the update is
    row-wise or a GEMM over the Bf*T rows."""
    threads = threads or usable_cpus()
    torch.set_num_threads(threads)
    t0 = time.perf_counter()
    t_step, parts = time_rollout_timestep(rollout_envs, 1, seed)
    t_step *= B / rollout_envs
    _log(f"{threads} threads; rollout timestep at {rollout_envs} envs: {t_step:.2f} s  " + " ".join(f"{k} {v:.2f}" for k, v in parts.items()))
    t_upd_s = time_update(update_episodes, seed)
    t_upd = t_upd_s * (B / update_episodes)
    _log(f"train() at Bf={update_episodes}: {t_upd_s:.2f} s -> {t_upd:.1f} s at Bf={B}")
    value = B * T / (T * t_step + t_upd)
    sample = (f"the reference's own code (oracle/_ref): ParallelRunner.run at {rollout_envs} envs, one timed loop iteration ("
              f"select_actions_ippo + GAT_latent_update + latent_update + EpisodeBatch.update) x{T} timesteps"
              f"{'' if rollout_envs == B else f' x{B / rollout_envs:g} envs'}; IPPOLearner.insert_episode_batch + train at Bf={update_episodes} "
              f"(T={T}, 15 epochs, 5 agents) x{B / update_episodes:g} in rows; {threads} torch threads")
    return dict(value=value, t_step=t_step, t_train=t_upd, parts=parts, cores=threads, sample=sample,
                sample_s=time.perf_counter() - t0)


# ---- BASELINE configs[4]: the reference's GAT_Net at hidden width 128 (synthetic GAT + GRU microbench) -----------------------
def time_gat_net_128(n_envs=16, n_nets=1, seed=0):
    """Seconds per (env, agent-net) item of the reference's own GAT_Net.forward (nova/GAT_Net.py:41-142) at
    GAT_hidden_dim = attention_dim = input width = 128 and 16 slots, measured on `n_nets` nets x `n_envs` envs."""
    activate()
    from nova.GAT_Net import GAT_Net
    args = ref_args("highway", GAT_hidden_dim=128, attention_dim=128)
    args.max_vehicle_num = 16
    torch.manual_seed(seed)
    nets = [GAT_Net(input_shape=128, args=args) for _ in range(n_nets)]
    x = torch.rand(n_envs, 16, 128) * 2 - 1
    h = torch.tanh(torch.randn(n_envs * 16, 128))
    with torch.no_grad():
        nets[0](x, h)                                                   # warm-up
        t0 = time.perf_counter()
        for net in nets:
            net(x, h)
        dt = time.perf_counter() - t0
    return dt / (n_envs * n_nets)
