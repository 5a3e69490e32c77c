"""Observation-history wrapper — host-side mirror of the reference's
``observersation_state_history_wrapper`` (/root/reference/observation_wrapper.py:6-141, sic) for the methods the
rollout runner calls every timestep (runners/ippo_parallel_runner.py:119-121, :218-219, :229):

    agent_obs_profile_init(obs)          reset at the start of an episode
    obs_history_create(obs)              assign vehicle ids to slots in first-seen order, append this step's rows
    obs_single_history_output()          np [B, A, N, o]      the newest row of every slot
    obs_history_output()                 np [B, A, N, W, o]   the last W rows of every slot, right aligned
    pure_obs_state_wrapper(state, obs)   strip the id column (numpy, not on the hot path)

The per-(env, agent) bookkeeping is kernel ``iplan_obs_history_step`` (csrc/obs_history.cu); the state (slot table,
windows) lives on the device, and ``.window`` / ``.single`` expose it as CUDA tensors so that K1 / K1b can read it in
place: ``window.view(B, A, N, W * o).permute(1, 0, 2, 3)`` is K1b's input layout.

Differences from the reference, by design: outputs are float32 (the reference builds float64 arrays that every consumer
casts to float32); a vehicle id repeated within one observation keeps its last row (the reference's deque would take two
appends in one timestep); meeting more than ``max_vehicle_num`` distinct ids raises (the reference fails with IndexError at
output time).

Beyond the reference's surface: ``step(obs_device)`` is the hook a vectorised (device-resident) simulator calls once per
timestep with its raw observation tensor — it runs the kernel and returns the ``(single, window)`` CUDA tensors, no host
round trip.  ``obs_history_episode_output(mask)`` (reference :145-173; unused by the reference's own training loop) is
served from a per-episode log of the appended rows kept on the device (``record_episode=True``).
"""
import numpy as np
import torch

from . import _lib


class observersation_state_history_wrapper:
    def __init__(self, args, n_agents, max_vehicle_num, max_episode_len, max_history_len):
        if not getattr(args, "use_cuda", True):
            raise RuntimeError("iplan_b200.observation_wrapper runs on CUDA only (no CPU path)")
        self.args = args
        self.max_vehicle_num = max_vehicle_num
        self.max_episode_len = max_episode_len
        self.max_history_len = max_history_len
        self.obs_shape = args.obs_shape_single
        self.n_agents = n_agents
        self.n_threads = args.batch_size_run
        self.device = torch.device("cuda")
        self.slot_ids = self.slot_count = self.window = self.single = self._overflow = None
        self.record_episode = True        # keep every appended row of the episode (for obs_history_episode_output)
        self.episode_log = None           # [B, A, N, max_episode_len + 1, o], column t = the rows appended at call t
        self.calls = 0

    def _alloc(self, B, A):
        N, W, o, dev = self.max_vehicle_num, self.max_history_len, self.obs_shape, self.device
        self.slot_ids = torch.full((B, A, N), -1, dtype=torch.int32, device=dev)
        self.slot_count = torch.zeros(B, A, dtype=torch.int32, device=dev)
        self.window = torch.zeros(B, A, N, W, o, device=dev)
        self.single = torch.zeros(B, A, N, o, device=dev)
        self._overflow = torch.zeros(1, dtype=torch.int32, device=dev)
        self.episode_log = torch.zeros(B, A, N, self.max_episode_len + 1, o, device=dev) if self.record_episode else None

    def agent_obs_profile_init(self, obs):
        """Reference :25-45.  ``obs`` [B, A, n_obs, obs_dim] (numpy or tensor); only its shape is used here."""
        B, A = int(obs.shape[0]), int(obs.shape[1])
        if self.window is None or tuple(self.window.shape[:2]) != (B, A):
            self._alloc(B, A)
        else:
            self.slot_ids.fill_(-1); self.slot_count.zero_(); self.window.zero_(); self.single.zero_(); self._overflow.zero_()
            if self.episode_log is not None:
                self.episode_log.zero_()
        self.calls = 0
        return self.slot_ids

    def step(self, obs_device):
        """Hook for a device-resident simulator: ``obs_device`` [B, A, n_obs, 1 + o] CUDA fp32 (column 0 = vehicle id).
        Runs the slot assignment + window update on the current stream and returns ``(single [B,A,N,o], window
        [B,A,N,W,o])`` — the tensors K1 / K1b read in place; nothing crosses PCIe."""
        x = obs_device
        assert torch.is_tensor(x) and x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()
        B, A, M, od = x.shape
        assert od == self.obs_shape + 1, (od, self.obs_shape)
        if self.window is None or tuple(self.window.shape[:2]) != (B, A):
            self.agent_obs_profile_init(x)
        _lib.check(_lib.lib.iplan_obs_history_step(
            _lib.ptr(x), B, A, M, od, _lib.ptr(self.slot_ids), _lib.ptr(self.slot_count), _lib.ptr(self.window),
            _lib.ptr(self.single), _lib.ptr(self._overflow), self.max_vehicle_num, self.max_history_len, _lib.stream()),
            "obs_history_step")
        if self.episode_log is not None and self.calls <= self.max_episode_len:
            self.episode_log[:, :, :, self.calls] = self.single
        self.calls += 1
        return self.single, self.window

    def obs_history_create(self, obs):
        """Reference :68-97.  Returns (agent ids [B, A] int tensor, slot table [B, A, N] int32, window tensor)."""
        x = obs if torch.is_tensor(obs) and obs.is_cuda else _lib.to_device(np.asarray(obs))
        x = x.to(torch.float32).contiguous()
        self.step(x)
        return x[:, :, 0, 0].to(torch.int64), self.slot_ids, self.window

    def obs_history_episode_output(self, mask):
        """Reference :145-173: the whole episode's rows of every slot, right aligned in ``max_episode_len`` columns and
        multiplied by ``mask[k, column, agent]``; returns (raw [B,A,N,L,o], the same reshaped [B,A,N,L/W,W,o])."""
        self._check()
        if self.episode_log is None:
            raise RuntimeError("obs_history_episode_output needs record_episode=True")
        L, W = self.max_episode_len, self.max_history_len
        B, A, N, _, o = self.episode_log.shape
        n = min(self.calls, L)                                  # a deque(maxlen=L) keeps the last L rows
        raw = torch.zeros(B, A, N, L, o, device=self.device)
        raw[:, :, :, L - n:] = self.episode_log[:, :, :, self.calls - n:self.calls]
        m = torch.as_tensor(np.asarray(mask), dtype=torch.float32, device=self.device)      # [B, L, A]
        raw = raw * m.permute(0, 2, 1)[:, :, None, :, None]
        # rows of a slot that did not exist yet are zero either way: the reference writes only len(history) columns
        raw_h = _lib.to_host(raw)
        return raw_h, raw_h.reshape(B, A, N, int(L // W), W, o)

    def _check(self):
        if int(self._overflow.item()):
            raise IndexError(f"an agent observed more than max_vehicle_num = {self.max_vehicle_num} distinct vehicles")

    def obs_history_output(self):
        self._check()
        return _lib.to_host(self.window)

    def obs_single_history_output(self):
        self._check()
        return _lib.to_host(self.single)

    def pure_obs_state_wrapper(self, state, obs):
        """Reference :51-59 (numpy slicing, unchanged)."""
        obs = np.asarray(obs)
        n_threads, n_agents, obs_num, obs_dim = obs.shape
        _, _, state_dim = state.shape
        n_vehicles = int(state_dim // obs_dim)
        new_state = state.reshape(n_threads, 1, n_vehicles, obs_dim)[:, :, :, 1:].reshape((n_threads, -1))
        new_obs = obs[:, :, :, 1:].reshape((n_threads, n_agents, -1))
        return new_state, new_obs
