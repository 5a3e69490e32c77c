"""Instant-incentive (GAT) module — host-side mirror of the reference's
``Prediction_policy`` (/root/reference/nova/prediction_policy.py:14-118) for the rollout
entry point ``GAT_latent_update``; the arithmetic is kernel K1 (csrc/gat_step.cu).

Same constructor, same ``GAT_latent_update(history_single, encoder_hidden,
behavior_latent) -> np.float32 [B, A, N, D]`` contract, same ``pred_GAT[i]`` modules
(state_dict keys/shapes) and ``pred_GAT_{i}.th`` checkpoint files.  The auxiliary
trajectory-prediction learner (``learn``, reference :168-253) is a "next" row of the
scope table and is not built.
"""
import copy
import os

import numpy as np
import torch

from .. import _lib
from ..modules.flat import ParamStack


class Prediction_policy:
    def __init__(self, args, logger=None):
        self.device = torch.device("cuda" if args.use_cuda else "cpu")
        if self.device.type != "cuda":
            raise RuntimeError("iplan_b200.Prediction_policy runs on CUDA only (no CPU path); set args.use_cuda=True")
        self.args = args
        self.logger = logger
        self.n_agents = args.n_agents
        self.max_vehicle_num = args.max_vehicle_num
        self.obs_shape = args.obs_shape_single
        if not getattr(args, "GAT_use_behavior", True):
            raise NotImplementedError("only the full iPLAN setting (GAT_use_behavior=True) is built")
        assert args.GAT_hidden_dim == 32 and args.attention_dim == 32, "kernel K1 is built for H = D = 32"
        self.GAT_input_dim = args.obs_shape_single + args.latent_dim          # reference :53-56
        self.stack = ParamStack("gat", self.n_agents, (self.GAT_input_dim,), device=self.device)
        self.pred_GAT = self.stack.nets
        self.tau = 0.01                                                       # nova/GAT_Net.py:93
        self.seed = int(getattr(args, "seed", 112358))
        self.calls = 0              # Philox counter: a fresh noise stream per call
        self.debug_gumbel = None    # [A,B,N,N-1,2] explicit gumbel noise for the next call (parity tests)
        self.capture_hard = False   # keep the hard-attention weights of the last call in .last_hard
        self.last_hard = None
        self._scratch = None      # kernel-to-kernel hand-off buffer of K1 (iplan_gat_scratch_floats)
        self._stage = None        # device staging buffers of the pipelined numpy entry point

    # ---- device path: tensors laid out [A, B, N, *] (any strides) ------------------
    def gat_step(self, hist, beh_prev, h_prev, out, gumbel=None, dbg_hard=None, events=None):
        """out[a,b,n,:] = GAT_a(hist[a,b], beh_prev[a,b], h_prev[a,b]); all CUDA fp32.
        ``events``: optional three recorded ``torch.cuda.Event(enable_timing=True)``; they are re-recorded on the
        launching stream before the recurrence kernel, between the two kernels and after the attention kernel."""
        A, B, N, o = hist.shape
        if gumbel is not None:
            assert gumbel.is_contiguous() and tuple(gumbel.shape) == (A, B, N, N - 1, 2), gumbel.shape
        need = _lib.lib.iplan_gat_scratch_floats(B, A, N)
        if self._scratch is None or self._scratch.numel() < need or self._scratch.device != hist.device:
            self._scratch = torch.empty(need, device=hist.device, dtype=torch.float32)
        ev = [None, None, None] if events is None else [e.cuda_event for e in events]
        rc = _lib.lib.iplan_gat_step_ex(
            _lib.ptr(self.stack.flat), self.stack.stride(),
            _lib.view(hist), _lib.view(beh_prev), _lib.view(h_prev), _lib.view(out),
            _lib.ptr(gumbel), self.seed, self.calls, self.tau, _lib.ptr(dbg_hard),
            _lib.ptr(self._scratch), self._scratch.numel(),
            B, A, N, o, beh_prev.shape[-1], ev[0], ev[1], ev[2], _lib.stream())
        _lib.check(rc, "gat_step")
        self.calls += 1
        return out

    # ---- reference-compatible numpy entry point (reference :92-118) -----------------
    def GAT_latent_update(self, history_single, encoder_hidden, behavior_latent=None):
        dev = self.device
        gum = self.debug_gumbel
        self.debug_gumbel = None
        perm = (1, 0, 2, 3)       # [B,A,N,*] -> [A,B,N,*] views, no copy
        hs_h, eh_h, bl_h = _lib.as_host(history_single), _lib.as_host(encoder_hidden), _lib.as_host(behavior_latent)
        B = hs_h.shape[0]
        if gum is None and not self.capture_hard and _lib.can_pipeline((hs_h, eh_h, bl_h), B):
            # page-locked inputs: copy-in, K1 and copy-out overlap chunk by chunk over the envs
            key = (tuple(hs_h.shape), tuple(eh_h.shape), tuple(bl_h.shape))
            if self._stage is None or self._stage[0] != key:
                self._stage = (key, torch.empty(hs_h.shape, device=dev), torch.empty(eh_h.shape, device=dev),
                               torch.empty(bl_h.shape, device=dev), torch.empty(eh_h.shape, device=dev))
            _, hs, eh, bl, out = self._stage
            out_h = torch.empty(eh_h.shape, dtype=torch.float32, pin_memory=True)

            def launch(lo, hi):
                self.gat_step(hs[lo:hi].permute(perm), bl[lo:hi].permute(perm), eh[lo:hi].permute(perm),
                              out[lo:hi].permute(perm))

            _lib.run_pipelined((hs_h, eh_h, bl_h), (hs, eh, bl), out_h, out, launch)
            return out_h.numpy()
        hs = _lib.to_device(hs_h)
        eh = _lib.to_device(eh_h)
        bl = _lib.to_device(bl_h)
        out = torch.empty_like(eh)
        dbg = None
        if self.capture_hard:
            B, A, N, _ = hs.shape
            dbg = torch.zeros(A, B, N, N - 1, device=dev)
            self.last_hard = dbg
        self.gat_step(hs.permute(perm), bl.permute(perm), eh.permute(perm), out.permute(perm), gum, dbg)
        return _lib.to_host(out)

    def learn(self, batch, t_env):
        raise NotImplementedError("Prediction_policy.learn (aux trajectory-prediction loss, reference "
                                  "nova/prediction_policy.py:168-253) is outside the built hot path (SURVEY §8f)")

    # ---- checkpoints (reference :256-285) ----------------------------------------
    def save_models(self, path):
        for i, net in enumerate(self.pred_GAT):
            torch.save({k: v.detach().cpu() for k, v in net.state_dict().items()}, f"{path}/pred_GAT_{i}.th")

    def load_models(self, paths, load_optimisers=False):
        if len(paths) == 1:
            paths = [copy.copy(paths[0]) for _ in range(self.n_agents)]
        for i, net in enumerate(self.pred_GAT):
            net.load_state_dict(torch.load(os.path.join(paths[i], f"pred_GAT_{i}.th"),
                                           map_location="cpu", weights_only=False))
