"""N>1 host logic on CPU: two gloo ranks, each holding half of the episodes, must reproduce
the single-rank quantities — advantage normalisation constants, loss denominators, the
"first batch_size episodes" rule, and the summed gradients (computed with the oracle's loss
on each shard with GLOBAL denominators, reduced through iplan_b200.parallel.GradBucket)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from iplan_b200 import parallel
from oracle import iplan_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make_data(seed=0, Bf=8, T=6):
    g = torch.Generator().manual_seed(seed)
    values = torch.randn(Bf, T + 1, generator=g)
    rewards = torch.randn(Bf, T, generator=g) * 2
    alive = torch.ones(Bf, T + 1)
    alive[1, 3:] = 0
    alive[6, 2:] = 0
    logp = -torch.rand(Bf, T, generator=g)
    old_logp = logp + 0.3 * torch.randn(Bf, T, generator=g)
    return values, rewards, alive, logp, old_logp


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        Bf, T = 8, 6
        values, rewards, alive, logp, old_logp = _make_data()
        batch_size_global = Bf - 1
        lo, hi = rank * Bf // world, (rank + 1) * Bf // world
        n_local = hi - lo
        n_train = parallel.shard_train_episodes(rank, world, n_local, batch_size_global)
        # --- per-shard GAE + raw advantages + moments (what gae_adv_kernel writes) --------------
        ret = O.gae_returns(values[lo:hi], rewards[lo:hi], alive[lo:hi])
        adv = ret - values[lo:hi, :T]
        adv[alive[lo:hi, :T] == 0] = 0
        mom = torch.tensor([[adv.double().sum(), (adv.double() ** 2).sum(), float(n_local * T),
                             alive[lo:lo + n_train, :T].double().sum()]], dtype=torch.float64)
        parallel.allreduce_sum_(mom)
        norm = parallel.advantage_norm_from_moments(mom, batch_size_global * T)[0]
        # --- per-shard policy-loss gradient wrt logp with GLOBAL denominators ------------------
        lp = logp[lo:hi].clone().requires_grad_(True)
        advn = (adv - norm[0]) * norm[1]
        ratio = torch.exp(lp - old_logp[lo:hi])
        s1, s2 = ratio * advn, torch.clamp(ratio, 0.8, 1.2) * advn
        w = torch.zeros(n_local, T)
        w[:n_train] = alive[lo:lo + n_train, :T]
        loss = (-torch.min(s1, s2) * w).sum() * norm[2]
        loss.backward()
        # a "parameter gradient" = shard-summed d loss / d logp projected on a fixed basis
        basis = torch.arange(1, T + 1, dtype=torch.float32)
        g1 = (lp.grad * basis).sum().view(1)
        g2 = lp.grad.sum().view(1)
        bucket = parallel.GradBucket()
        bucket.allreduce([g1, g2])
        total_loss = loss.detach().clone().view(1)
        parallel.allreduce_sum_(total_loss)
        if rank == 0:
            out.put(dict(norm=norm, g1=g1, g2=g2, loss=total_loss, n_train=n_train))
        else:
            out.put(dict(n_train=n_train))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_reproduces_single_rank():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    r0 = [r for r in res if "norm" in r][0]
    assert sorted(r["n_train"] for r in res) == [3, 4]          # 7 of 8 episodes: last rank drops one

    # single-rank reference (the oracle's restatement of learners/ippo_learner.py:270-279, :185-197)
    Bf, T = 8, 6
    values, rewards, alive, logp, old_logp = _make_data()
    ret = O.gae_returns(values, rewards, alive)
    advn = O.normalised_advantages(ret, values[:, :T], alive[:, :T])
    adv = ret - values[:, :T]
    adv[alive[:, :T] == 0] = 0
    std, mean = torch.std_mean(adv)
    assert abs(float(r0["norm"][0]) - float(mean)) < 1e-6
    assert abs(float(r0["norm"][1]) - 1.0 / (float(std) + 1e-5)) < 1e-5
    nb = Bf - 1
    lp = logp.clone().requires_grad_(True)
    loss, _ = O.policy_loss_terms(lp[:nb].reshape(-1), old_logp[:nb].reshape(-1), advn[:nb].reshape(-1),
                                  alive[:nb, :T].reshape(-1))
    loss.backward()
    basis = torch.arange(1, T + 1, dtype=torch.float32)
    assert abs(float(r0["loss"]) - float(loss)) < 1e-5
    assert abs(float(r0["g1"]) - float((lp.grad * basis).sum())) < 1e-5
    assert abs(float(r0["g2"]) - float(lp.grad.sum())) < 1e-5


def test_shard_plan():
    assert [parallel.shard_train_episodes(r, 8, 512, 4095) for r in range(8)] == [512] * 7 + [511]
    assert parallel.shard_train_episodes(0, 1, 512, 511) == 511
    assert [parallel.shard_train_episodes(r, 4, 4, 9) for r in range(4)] == [4, 4, 1, 0]
