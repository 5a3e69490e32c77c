"""Pin the oracle (oracle/iplan_oracle.py) against outputs of the reference's own
modules (tests/golden/*.pt, written by tests/golden/make_golden.py in the build
container).  CPU only."""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import iplan_oracle as O

TOL = 2e-6


def load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name), weights_only=False)


def maxdiff(a, b):
    a = torch.as_tensor(np.asarray(a) if not torch.is_tensor(a) else a).double()
    b = torch.as_tensor(np.asarray(b) if not torch.is_tensor(b) else b).double()
    return float((a - b).abs().max())


@pytest.mark.parametrize("case", ["mpe", "highway"])
def test_gat_forward_matches_reference(golden_dir, case):
    g = load(golden_dir, "gat_net.pt")[case]
    out = O.gat_forward(g["params"], g["obs"], g["h_prev"], g["gumbel"])
    assert maxdiff(out, g["out"]) < 5e-6
    out2 = O.gat_forward_loops(g["params"], g["obs"], g["h_prev"], g["gumbel"])
    assert maxdiff(out2, g["out"]) < 5e-6


@pytest.mark.parametrize("case", ["mpe", "highway"])
def test_rollout_modules_match_reference(golden_dir, case):
    g = load(golden_dir, "rollout_modules.pt")[case]
    A = len(g["gat"])
    avail = torch.as_tensor(g["avail"])
    for t, st in enumerate(g["steps"]):
        att = O.gat_latent_update(g["gat"], st["history_single"], st["att_in"], st["beh_in"], st["gumbel"])
        assert maxdiff(att, st["att_out"]) < 5e-6, t
        beh_now = torch.as_tensor(st["beh_in"])
        if "beh_out" in st:
            beh_now, hid = O.behavior_latent_update(g["beh"], st["window"], st["enc_rnn_in"], st["beh_in"])
            assert maxdiff(beh_now, st["beh_out"]) < TOL, t
            assert maxdiff(hid, st["enc_rnn_out"]) < TOL, t
        B = att.shape[0]
        if t == 0:
            last = torch.zeros(B, A, avail.shape[-1])
        else:
            last = torch.nn.functional.one_hot(torch.as_tensor(g["steps"][t - 1]["actions"]),
                                               avail.shape[-1]).float()
        x = O.build_inputs_step(torch.as_tensor(st["history_single"], dtype=torch.float32),
                                torch.as_tensor(st["att_out"]), beh_now.float(), last, A)
        assert maxdiff(x, st["inputs"]) < TOL, t
        r = O.select_actions(g["actors"], g["critics"], x, avail,
                             torch.as_tensor(st["rnn_a_in"]).reshape(B, A, -1),   # [B,1,A,R] at t=0, [1,B,A,R] after
                             torch.as_tensor(st["rnn_c_in"]).reshape(B, A, -1),
                             test_mode=True)
        assert maxdiff(r["values"], st["values"]) < 2e-5, t
        assert (r["actions"].numpy() == st["actions"]).all(), t
        assert maxdiff(r["logp"], st["logp"].reshape(B, A)) < 2e-5, t
        assert maxdiff(r["rnn_a"], np.asarray(st["rnn_a_out"])[0]) < TOL * 5, t
        assert maxdiff(r["rnn_c"], np.asarray(st["rnn_c_out"])[0]) < TOL * 5, t


@pytest.mark.parametrize("case", ["mpe", "highway"])
def test_learner_matches_reference(golden_dir, case):
    g = load(golden_dir, "learner.pt")[case]
    args = SimpleNamespace(**g["args"])
    d = g["data"]
    A = args.n_agents
    onehot = torch.nn.functional.one_hot(d["actions"].squeeze(-1), args.n_actions).float()
    all_stats = []
    for a in range(A):
        batch = dict(
            history=d["history"][:, :, a], attention_latent=d["attention_latent"][:, :, a],
            behavior_latent=d["behavior_latent"][:, :, a], actions=d["actions"][:, :, a],
            actions_onehot=onehot[:, :, a], available_actions=d["avail_actions"][:, :, a],
            reward=d["reward"][:, :, a], terminated_masks=(1 - d["terminated"][:, :, a].float()),
            rnn_states_actor=d["rnn_states_actors"][:, :, a], rnn_states_critic=d["rnn_states_critics"][:, :, a])
        ap = {k: v.clone() for k, v in g["actors_before"][a].items()}
        cp = {k: v.clone() for k, v in g["critics_before"][a].items()}
        obs_all = O.build_inputs_train(a, batch["history"], batch["attention_latent"],
                                       batch["behavior_latent"], batch["actions_onehot"], A)
        assert maxdiff(obs_all, g["pre"][a]["obs_all"]) < TOL
        stats, pre, _, _ = O.train_agent(ap, cp, batch, a, args, perms=g["perms"][a])
        assert maxdiff(pre["values_all"], g["pre"][a]["values_all"]) < 2e-5
        assert maxdiff(pre["returns"], g["pre"][a]["returns"]) < 5e-5
        assert maxdiff(pre["advantages"], g["pre"][a]["advantages"]) < 5e-5
        assert maxdiff(pre["old_logp"], g["pre"][a]["old_logp"]) < 2e-5
        for k in O.ACTOR_TRAINABLE:
            assert maxdiff(ap[k], g["actors_after"][a][k]) < 2e-5, (a, k)
        for k in O.CRITIC_TRAINABLE:
            assert maxdiff(cp[k], g["critics_after"][a][k]) < 2e-5, (a, k)
        # fc_h is in the state_dict but never trained (mlp.py:20-27)
        for k in g["actors_before"][a]:
            if "fc_h" in k:
                assert torch.equal(g["actors_before"][a][k], g["actors_after"][a][k])
        all_stats += stats
    n_upd = len(all_stats)
    for key in ("value_loss", "policy_loss", "dist_entropy", "actor_grad_norm", "critic_grad_norm", "ratio"):
        mine = sum(s[key] for s in all_stats) / n_upd
        ref = [v for k, v in g["stats"].items() if k.endswith(key)][0]
        assert abs(mine - ref) < 1e-4 * max(1.0, abs(ref)), (key, mine, ref)
    assert g["actor_opt_steps"][0] == len(O.ACTOR_TRAINABLE)   # 18 of 22 tensors have Adam state


def test_one_sided_huber():
    e = torch.tensor([-20.0, -5.0, 0.0, 5.0, 20.0])
    assert O.huber_one_sided(e, 10.0).tolist() == [0.0, 12.5, 0.0, 12.5, 150.0]


@pytest.mark.parametrize("case", ["mpe", "highway"])
def test_prediction_learn_matches_reference(golden_dir, case):
    """SURVEY §8f rank 2 — Prediction_policy.learn: the oracle (autograd through the oracle's own GAT forward, the
    restated decoder, the masked L1 loss, two-group gradient clipping, one Adam) against one call of the reference's
    ``learn`` with its random draws recorded (tests/golden/make_golden.py::golden_prediction_learn)."""
    from types import SimpleNamespace
    g = torch.load(os.path.join(golden_dir, "prediction_learn.pt"), weights_only=False)[case]
    args = SimpleNamespace(**g["args"])
    d = g["data"]
    A = args.n_agents
    hist, att, beh = d["history"][:, :-1], d["attention_latent"][:, :-1], d["behavior_latent"][:, :-1]
    flag = d["terminated"][:, :-1, :, 0]
    for a in range(A):
        gp = {k: v.clone() for k, v in g["gat_before"][a].items()}
        dp = {k: v.clone() for k, v in g["dec_before"][a].items()}
        out, _ = O.prediction_learn_agent(gp, dp, hist[:, :, a], att[:, :, a], beh[:, :, a], flag[:, :, a],
                                          g["select_idx"][a], g["gumbel"][a], g["dropout_keep"][a], args)
        assert abs(out["loss"] - g["losses"][a]) < 1e-5 * max(1.0, abs(g["losses"][a])), (a, out["loss"], g["losses"][a])
        worst = max(float((gp[k] - g["gat_after"][a][k]).abs().max()) for k in gp)
        worst = max(worst, max(float((dp[k] - g["dec_after"][a][k]).abs().max()) for k in dp))
        moved = max(float((g["gat_after"][a][k] - g["gat_before"][a][k]).abs().max()) for k in gp)
        print(f"[prediction.learn {case} a={a}] loss {out['loss']:.6f} (reference {g['losses'][a]:.6f}); "
              f"max |param - reference| after the step {worst:.2e} (the step moved them by {moved:.2e})")
        assert worst < 2e-7 and moved > 1e-6
        # Adam's first step is ~lr * sign(g): compare the gradients themselves (as clipped by the reference)
        ref_g = {**g["gat_grads"][a], **g["dec_grads"][a]}
        rel = max(float((out["clipped"][k] - ref_g[k]).abs().max() / (ref_g[k].abs().max() + 1e-12)) for k in ref_g)
        print(f"[prediction.learn {case} a={a}] worst relative gradient difference {rel:.2e}; "
              f"norms GAT {out['gat_grad_norm']:.4f} decoder {out['dec_grad_norm']:.4f}")
        assert rel < 2e-4
    stats = g["stats"]
    key = [k for k in stats if k.endswith("prediction_loss")][0]
    assert abs(stats[key] - sum(g["losses"])) < 1e-4 * abs(stats[key])


@pytest.mark.parametrize("case", ["mpe", "highway"])
def test_behavior_learn_matches_reference(golden_dir, case):
    """SURVEY §8f rank 3 — Behavior_policy.learn against one call of the reference's ``learn`` with its dropout draws
    recorded: losses, the gradients as clipped by the reference (BPTT across every window position), post-step weights."""
    from types import SimpleNamespace
    g = torch.load(os.path.join(golden_dir, "behavior_learn.pt"), weights_only=False)[case]
    args = SimpleNamespace(**g["args"])
    d = g["data"]
    hist = d["history"][:, :-1]
    term = d["terminated"][:, :-1, :, 0].float()
    for a in range(args.n_agents):
        mask = 1 - term[:, :, a] if args.env == "MPE" else term[:, :, a]
        ep = {k: v.clone() for k, v in g["enc_before"][a].items()}
        dp = {k: v.clone() for k, v in g["dec_before"][a].items()}
        out, _ = O.behavior_learn_agent(ep, dp, hist[:, :, a], mask, g["dropout_keep"][a], args)
        assert abs(out["behavior_loss"] - g["behavior_loss"][a]) < 2e-5 * abs(g["behavior_loss"][a])
        assert abs(out["stability_loss"] - g["stability_loss"][a]) < 2e-5 * abs(g["stability_loss"][a])
        ref_g = {**{"enc:" + k: v for k, v in g["enc_grads"][a].items()}, **{"dec:" + k: v for k, v in g["dec_grads"][a].items()}}
        rel = max(float((out["clipped"][k] - ref_g[k]).abs().max() / (ref_g[k].abs().max() + 1e-12)) for k in ref_g)
        worst = max(max(float((ep[k] - g["enc_after"][a][k]).abs().max()) for k in ep),
                    max(float((dp[k] - g["dec_after"][a][k]).abs().max()) for k in dp))
        print(f"[behavior.learn {case} a={a}] loss {out['behavior_loss']:.6f} (reference {g['behavior_loss'][a]:.6f}); "
              f"worst relative gradient difference {rel:.2e}; max |param - reference| after the step {worst:.2e}")
        assert rel < 5e-4 and worst < 1e-6


def test_obs_history_matches_reference(golden_dir):
    """SURVEY §8f rank 4 — the observation-history wrapper restated as one shift-and-append over the slots, against
    14 timesteps recorded from the reference's own wrapper (vehicles enter, leave and re-enter the view)."""
    g = torch.load(os.path.join(golden_dir, "obs_wrapper.pt"), weights_only=False)
    d = g["dims"]
    oh = O.ObsHistory(d["B"], d["A"], d["N"], d["W"], d["o"])
    for t, st in enumerate(g["steps"]):
        win, single = oh.step(st["obs"].numpy())
        assert (win == st["window"].numpy()).all(), t
        assert (single == st["single"].numpy()).all(), t
    assert oh.ids == g["ids"]


@pytest.mark.parametrize("case", ["mpe", "highway"])
def test_gat_backward_by_hand_matches_autograd(golden_dir, case):
    """The hand-written backward of GAT_Net (explicit BPTT over the bidirectional hard-attention GRU, the gumbel-sigmoid
    gate, soft attention, GRUCell) — the arithmetic specification of the K1 backward kernels — against autograd through
    the oracle's forward, which the prediction-learn fixture pins to the reference."""
    g = torch.load(os.path.join(golden_dir, "gat_net.pt"), weights_only=False)[case]
    p = {k: v.clone().requires_grad_(True) for k, v in g["params"].items()}
    torch.manual_seed(3)
    d_out = torch.randn_like(g["out"])
    out = O.gat_forward(p, g["obs"], g["h_prev"], g["gumbel"])
    auto = torch.autograd.grad((out * d_out).sum(), list(p.values()))
    hand = O.gat_backward_manual({k: v.detach() for k, v in p.items()}, g["obs"], g["h_prev"], g["gumbel"], d_out)
    for (k, _), ga in zip(p.items(), auto):
        rel = float((hand[k] - ga).abs().max() / (ga.abs().max() + 1e-12))
        assert rel < 5e-4, (k, rel)
