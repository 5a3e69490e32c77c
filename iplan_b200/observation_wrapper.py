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
output time).  ``obs_history_episode_output`` (used only by the auxiliary learners) is not built.
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

    def _alloc(self, B, A):
        N, W, o, dev = self.max_vehicle_num, self.max_history_len, self.obs_shape, self.device
        self.slot_ids = torch.full((B, A, N), -1, dtype=torch.int32, device=dev)
        self.slot_count = torch.zeros(B, A, dtype=torch.int32, device=dev)
        self.window = torch.zeros(B, A, N, W, o, device=dev)
        self.single = torch.zeros(B, A, N, o, device=dev)
        self._overflow = torch.zeros(1, dtype=torch.int32, device=dev)

    def agent_obs_profile_init(self, obs):
        """Reference :25-45.  ``obs`` [B, A, n_obs, obs_dim] (numpy or tensor); only its shape is used here."""
        B, A = int(obs.shape[0]), int(obs.shape[1])
        if self.window is None or tuple(self.window.shape[:2]) != (B, A):
            self._alloc(B, A)
        else:
            self.slot_ids.fill_(-1); self.slot_count.zero_(); self.window.zero_(); self.single.zero_(); self._overflow.zero_()
        return self.slot_ids

    def obs_history_create(self, obs):
        """Reference :68-97.  Returns (agent ids [B, A] int tensor, slot table [B, A, N] int32, window tensor)."""
        x = obs if torch.is_tensor(obs) and obs.is_cuda else _lib.to_device(np.asarray(obs))
        x = x.to(torch.float32).contiguous()
        B, A, M, od = x.shape
        assert od == self.obs_shape + 1, (od, self.obs_shape)
        _lib.check(_lib.lib.iplan_obs_history_step(
            _lib.ptr(x), B, A, M, od, _lib.ptr(self.slot_ids), _lib.ptr(self.slot_count), _lib.ptr(self.window),
            _lib.ptr(self.single), _lib.ptr(self._overflow), self.max_vehicle_num, self.max_history_len, _lib.stream()),
            "obs_history_step")
        return x[:, :, 0, 0].to(torch.int64), self.slot_ids, self.window

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
