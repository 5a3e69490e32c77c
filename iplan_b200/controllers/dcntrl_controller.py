"""Non-parameter-sharing multi-agent controller — host-side mirror of the reference's
``DcntrlMAC`` (/root/reference/controllers/dcntrl_controller.py:9-233).  The arithmetic is
kernel K1c (csrc/controller_step.cu) for the rollout step and the learner kernels
(csrc/learner.cu) for ``get_value_ippo`` / ``eval_action_ippo``.

Same constructor ``DcntrlMAC(scheme, groups, args)``, same public methods and return
types (numpy where the reference returns numpy), same ``agents[i]`` / ``critics[i]``
state_dict keys and ``agent_{i}.th`` / ``critic_{i}.th`` checkpoint files.
"""
import copy
import os

import numpy as np
import torch as th

from .. import _lib
from ..modules.flat import ParamStack


class DcntrlMAC:
    def __init__(self, scheme, groups, args):
        self.n_agents = args.n_agents
        self.args = args
        if not args.use_cuda:
            raise RuntimeError("iplan_b200.DcntrlMAC runs on CUDA only (no CPU path); set args.use_cuda=True")
        for flag in ("GAT_enable", "Behavior_enable", "obs_last_action", "obs_agent_id"):
            if not getattr(args, flag):
                raise NotImplementedError(f"only the full iPLAN setting is built ({flag}=True)")
        assert args.rnn_hidden_dim == 64 and args.mlp_hidden_dim == 64 and args.recurrent_N == 1 and args.layer_N == 1
        self.device = th.device("cuda")
        self.input_shape = self._get_input_shape(scheme)
        self.actor_stack = ParamStack("actor", self.n_agents, (self.input_shape, args.n_actions), device=self.device)
        self.critic_stack = ParamStack("critic", self.n_agents, (self.input_shape,), device=self.device)
        self.agents = self.actor_stack.nets
        self.critics = self.critic_stack.nets
        self.agent_output_type = getattr(args, "agent_output_type", None)
        self.action_selector = None      # epsilon-greedy selector is constructed but unused by IPPO (:21)
        self.hidden_states = None
        self.input_scheme = scheme
        self.seed = int(getattr(args, "seed", 112358))
        self.calls = 0
        self.debug_uniforms = None       # [A,B] explicit sampling noise for the next call (parity tests)
        self.last_logits = None
        self.capture_logits = False

    # ---- K1c on device tensors ------------------------------------------------------
    def controller_step(self, feat, rnn_a_in, rnn_c_in, rnn_a_out, rnn_c_out, avail_u8=None,
                        test_mode=False, uniforms=None, next_onehot=None, this_onehot=None, logits=None, out=None):
        """feat [A,B,>=F] rows (strided), rnn_* [A,B,64] (strided views).  Returns
        (actions int32 [A,B], logp [A,B], values [A,B]); ``out`` = caller-owned contiguous tensors for the three
        (a rollout loop passes slices of per-episode buffers: no allocation, no copy per timestep)."""
        A, B = feat.shape[0], feat.shape[1]
        assert feat.stride(2) == 1 and rnn_a_in.stride(2) == 1 and rnn_a_out.stride(2) == 1
        assert rnn_a_in.stride() == rnn_c_in.stride() and rnn_a_out.stride() == rnn_c_out.stride()
        if out is not None:
            actions, logp, values = out
            assert actions.is_contiguous() and logp.is_contiguous() and values.is_contiguous() and actions.dtype == th.int32
        else:
            actions = th.empty(A, B, dtype=th.int32, device=self.device)
            logp = th.empty(A, B, device=self.device)
            values = th.empty(A, B, device=self.device)
        rc = _lib.lib.iplan_controller_step(
            _lib.ptr(self.actor_stack.flat), self.actor_stack.stride(),
            _lib.ptr(self.critic_stack.flat), self.critic_stack.stride(),
            _lib.ptr(feat), feat.stride(0), feat.stride(1),
            _lib.ptr(rnn_a_in), _lib.ptr(rnn_c_in), _lib.ptr(rnn_a_out), _lib.ptr(rnn_c_out),
            rnn_a_in.stride(0), rnn_a_in.stride(1), rnn_a_out.stride(0), rnn_a_out.stride(1),
            _lib.ptr(avail_u8), _lib.ptr(uniforms), self.seed, self.calls, 1 if test_mode else 0,
            _lib.ptr(actions), _lib.ptr(logp), _lib.ptr(values), _lib.ptr(logits),
            _lib.ptr(next_onehot), _lib.ptr(this_onehot),
            B, A, self.input_shape, self.args.n_actions, _lib.stream())
        _lib.check(rc, "controller_step")
        self.calls += 1
        return actions, logp, values

    # ---- IPPO rollout entry point (reference :27-58) ---------------------------------
    def select_actions_ippo(self, ep_batch, t_ep, test_mode=False):
        A, R = self.n_agents, self.args.rnn_hidden_dim
        packed = getattr(ep_batch, "packed", None)
        if packed is not None:
            feat = packed[:, :, t_ep]                                    # [A,B,Fp] view, zero copy
        else:
            feat = self._build_inputs(ep_batch, t_ep).to(self.device, th.float32).permute(1, 0, 2).contiguous()
        B = feat.shape[1]
        rnn_a = ep_batch["rnn_states_actors"][:, t_ep].to(self.device).permute(1, 0, 2)     # [A,B,R] views
        rnn_c = ep_batch["rnn_states_critics"][:, t_ep].to(self.device).permute(1, 0, 2)
        avail = ep_batch["avail_actions"][:, t_ep].to(self.device)
        avail_u8 = (avail != 0).permute(1, 0, 2).contiguous().to(th.uint8)
        new_a = th.empty(A, B, R, device=self.device)
        new_c = th.empty(A, B, R, device=self.device)
        logits = th.empty(A, B, self.args.n_actions, device=self.device) if self.capture_logits else None
        uni = self.debug_uniforms
        self.debug_uniforms = None
        actions, logp, values = self.controller_step(feat, rnn_a, rnn_c, new_a, new_c, avail_u8,
                                                     test_mode=test_mode, uniforms=uni, logits=logits)
        self.last_logits = logits
        action_log_probs = [logp[a].view(B, 1) for a in range(A)]                   # list[A] of [B,1]
        # four small results, one stream synchronize: values [B,A], actions [B,A] int64, rnn states [1,B,A,R] (handed back
        # via batch.update: they keep a device shadow)
        values_np, actions_np, rnn_a_np, rnn_c_np = _lib.to_host_many(
            [values.t(), actions.t().to(th.int64), new_a.permute(1, 0, 2).unsqueeze(0), new_c.permute(1, 0, 2).unsqueeze(0)], shadows=(2, 3))
        return values_np, actions_np, action_log_probs, rnn_a_np, rnn_c_np

    # ---- learner-facing evaluation helpers (reference :61-85); bound by the learner ----
    def get_value_ippo(self, agent_id, obs, rnn_states_critic):
        from ..learners.ippo_learner import eval_rows
        v, _, _ = eval_rows(self, agent_id, obs, rnn_states_critic, net="critic")
        return v.reshape(*obs.shape[:-1], 1)

    def eval_action_ippo(self, agent_id, obs, action, available_actions, rnn_states_actor):
        from ..learners.ippo_learner import eval_rows
        _, logp, ent = eval_rows(self, agent_id, obs, rnn_states_actor, net="actor", action=action,
                                 avail=available_actions)
        return logp.reshape(*obs.shape[:-1], 1), ent

    # ---- input assembly (reference :87-115, :187-213) -- torch ops, used off the hot path
    def _slots(self, batch):
        return th.cat([batch["history"], batch["attention_latent"], batch["behavior_latent"]], dim=-1)

    def _build_inputs_ippo(self, agent_id, batch, action_onehot, discr_signal=None):
        bs, num_ts = batch["history"].shape[:2]
        slots = self._slots(batch).reshape(bs, num_ts, -1)
        last = th.cat([action_onehot[:, 0].unsqueeze(1), action_onehot[:, :-1]], dim=1)
        ident = th.zeros((bs, num_ts, self.n_agents), device=slots.device)
        ident[:, :, agent_id] = 1
        return th.cat([slots, last.to(slots.device), ident], dim=-1)

    def _build_inputs(self, batch, t):
        bs = batch.batch_size
        slots = th.cat([batch["history"][:, t], batch["attention_latent"][:, t], batch["behavior_latent"][:, t]], dim=-1)
        last = th.zeros_like(batch["actions_onehot"][:, t]) if t == 0 else batch["actions_onehot"][:, t - 1]
        eye = th.eye(self.n_agents, device=slots.device).unsqueeze(0).expand(bs, -1, -1)
        return th.cat([x.reshape(bs, self.n_agents, -1) for x in (slots, last.to(slots.device), eye)], dim=2)

    def _get_input_shape(self, scheme):
        n, o = scheme["history"]["vshape"]
        shape = n * o
        shape += scheme["attention_latent"]["vshape"][0] * scheme["attention_latent"]["vshape"][1]
        shape += scheme["behavior_latent"]["vshape"][0] * scheme["behavior_latent"]["vshape"][1]
        shape += scheme["actions_onehot"]["vshape"][0]
        shape += self.n_agents
        return shape

    # ---- housekeeping, same names as the reference (:117-174) ---------------------------
    def init_hidden(self, batch_size):
        self.hidden_states = [th.zeros(batch_size, 1, self.args.rnn_hidden_dim, device=self.device)
                              for _ in range(self.n_agents)]

    def parameters(self):
        return [list(agent.parameters()) for agent in self.agents]

    def critic_parameters(self):
        return [list(critic.parameters()) for critic in self.critics]

    def load_state(self, other_mac):
        for i, agent in enumerate(self.agents):
            agent.load_state_dict(other_mac.agents[i].state_dict())

    def cuda(self):
        self.actor_stack.to("cuda")
        self.critic_stack.to("cuda")

    def set_train_mode(self):
        pass        # no dropout / batch-norm on this path: train and eval are the same function

    def set_eval_mode(self):
        pass

    def save_models(self, path):
        for i, agent in enumerate(self.agents):
            th.save({k: v.detach().cpu() for k, v in agent.state_dict().items()}, f"{path}/agent_{i}.th")
        quirk = bool(getattr(self.args, "popart_cuda_quirk", False))      # CUDA reference: no v_out.* keys (modules/flat.py)
        for i, critic in enumerate(self.critics):
            th.save({k: v.detach().cpu() for k, v in critic.state_dict().items() if not (quirk and k.startswith("v_out."))},
                    f"{path}/critic_{i}.th")

    def load_models(self, paths):
        if len(paths) == 1:
            paths = [copy.copy(paths[0]) for _ in range(self.n_agents)]
        for i, agent in enumerate(self.agents):
            agent.load_state_dict(th.load(os.path.join(paths[i], f"agent_{i}.th"), map_location="cpu", weights_only=False))
        for i, critic in enumerate(self.critics):
            critic.load_state_dict(th.load(os.path.join(paths[i], f"critic_{i}.th"), map_location="cpu", weights_only=False))
