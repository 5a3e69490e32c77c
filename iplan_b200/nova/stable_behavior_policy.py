"""Behavioural-incentive module (soft-update variant = iPLAN) — host-side mirror of the
reference's ``Behavior_policy`` (/root/reference/nova/stable_behavior_policy.py:13-123) for
the rollout entry point ``latent_update``; the arithmetic is kernel K1b
(csrc/behavior_step.cu).

Same constructor, same ``latent_update(history, encoder_hidden, prev_latent) ->
(np latent [B,A,N,L], torch hidden [B,1,A,N,E])`` contract (the reference returns the
hidden state as a torch tensor and accepts numpy on the first call / a tensor
afterwards, :101-121), same ``behavior_encoder[i]`` state_dict keys and
``behavior_encoder_{i}.th`` files.  The auxiliary reconstruction learner (``learn``,
reference :161-279) is a "next" row of the scope table and is not built.
"""
import copy
import os

import numpy as np
import torch

from .. import _lib
from ..modules.flat import ParamStack


class Behavior_policy:
    def __init__(self, args, logger=None):
        self.device = torch.device("cuda" if args.use_cuda else "cpu")
        if self.device.type != "cuda":
            raise RuntimeError("iplan_b200.Behavior_policy runs on CUDA only (no CPU path); set args.use_cuda=True")
        self.args = args
        self.logger = logger
        self.n_agents = args.n_agents
        self.max_vehicle_num = args.max_vehicle_num
        self.max_history_len = args.max_history_len
        self.latent_dim = args.latent_dim
        self.soft_update_coef = args.soft_update_coef
        assert args.encoder_rnn_dim == 32 and args.num_encoder_layer == 1, "kernel K1b is built for E = 32, one layer"
        self.stack = ParamStack("beh", self.n_agents, (args.obs_shape_single, args.latent_dim), device=self.device)
        self.behavior_encoder = self.stack.nets
        # the reconstruction decoder of the auxiliary learner (reference :48-53): parameters and checkpoints only — ``learn`` is not built
        self.dec_stack = ParamStack("bdec", self.n_agents, (args.obs_shape_single, args.latent_dim, args.decoder_rnn_dim), device=self.device)
        self.behavior_decoder = self.dec_stack.nets
        self._stage = None        # device staging buffers of the pipelined numpy entry point

    # ---- device path: tensors laid out [A, B, N, *] ----------------------------------
    def behavior_step(self, window, hid_io, lat_prev, lat_out):
        """window [A,B,N,W*o], hid_io [A,B,N,E] (in place), lat_prev/lat_out [A,B,N,L]."""
        A, B, N, _ = window.shape
        rc = _lib.lib.iplan_behavior_step(
            _lib.ptr(self.stack.flat), self.stack.stride(),
            _lib.view(window), _lib.view(hid_io), _lib.view(lat_prev), _lib.view(lat_out),
            float(self.soft_update_coef), B, A, N, self.args.obs_shape_single, self.latent_dim,
            self.max_history_len, _lib.stream())
        _lib.check(rc, "behavior_step")
        return lat_out, hid_io

    # ---- reference-compatible entry point (reference :83-123) -------------------------
    def latent_update(self, history, encoder_hidden, prev_latent):
        dev = self.device
        hist_h, prev_h = _lib.as_host(history), _lib.as_host(prev_latent)
        B, A, N, W, o = hist_h.shape
        if torch.is_tensor(encoder_hidden) and encoder_hidden.is_cuda:
            hid = encoder_hidden.detach().to(torch.float32).clone()
        else:
            hid = _lib.to_device(encoder_hidden)
        perm = (1, 0, 2, 3)
        if _lib.can_pipeline((hist_h, prev_h), B):
            # page-locked inputs: copy-in, K1b and copy-out overlap chunk by chunk over the envs
            key = (tuple(hist_h.shape), tuple(prev_h.shape))
            if self._stage is None or self._stage[0] != key:
                self._stage = (key, torch.empty(hist_h.shape, device=dev), torch.empty(prev_h.shape, device=dev),
                               torch.empty(prev_h.shape, device=dev))
            _, hist, prev, new = self._stage
            new_h = torch.empty(prev_h.shape, dtype=torch.float32, pin_memory=True)

            def launch(lo, hi):
                self.behavior_step(hist[lo:hi].reshape(hi - lo, A, N, W * o).permute(perm), hid[lo:hi, 0].permute(perm),
                                   prev[lo:hi].permute(perm), new[lo:hi].permute(perm))

            _lib.run_pipelined((hist_h, prev_h), (hist, prev), new_h, new, launch)
            return new_h.numpy(), hid
        hist = _lib.to_device(hist_h)
        prev = _lib.to_device(prev_h)
        new = torch.empty_like(prev)
        hid_v = hid[:, 0].permute(perm)                          # [B,1,A,N,E] -> [A,B,N,E] view
        self.behavior_step(hist.reshape(B, A, N, W * o).permute(perm), hid_v,
                           prev.permute(perm), new.permute(perm))
        return _lib.to_host(new), hid

    def learn(self, batch, t_env):
        raise NotImplementedError("Behavior_policy.learn (aux reconstruction loss, reference "
                                  "nova/stable_behavior_policy.py:161-279) is outside the built hot path (SURVEY §8f)")

    # ---- checkpoints (reference :282-312) -------------------------------------------
    def save_models(self, path):
        for i, net in enumerate(self.behavior_encoder):
            torch.save({k: v.detach().cpu() for k, v in net.state_dict().items()}, f"{path}/behavior_encoder_{i}.th")
        for i, net in enumerate(self.behavior_decoder):
            torch.save({k: v.detach().cpu() for k, v in net.state_dict().items()}, f"{path}/behavior_decoder_{i}.th")

    def load_models(self, paths, load_optimisers=False):
        if len(paths) == 1:
            paths = [copy.copy(paths[0]) for _ in range(self.n_agents)]
        for i, net in enumerate(self.behavior_encoder):
            net.load_state_dict(torch.load(os.path.join(paths[i], f"behavior_encoder_{i}.th"),
                                           map_location="cpu", weights_only=False))
        for i, net in enumerate(self.behavior_decoder):
            f = os.path.join(paths[i], f"behavior_decoder_{i}.th")
            if os.path.exists(f):
                net.load_state_dict(torch.load(f, map_location="cpu", weights_only=False))
