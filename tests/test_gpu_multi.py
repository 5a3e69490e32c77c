"""Multi-GPU equivalence on real devices (needs >= 2 GPUs; skipped otherwise): two NCCL ranks,
each holding half of the episodes, must end a train() with the same weights as one rank
holding all of them (global advantage moments / mask sums / first-batch_size rule + one
gradient all-reduce per PPO epoch)."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _data(B, T, A, N, seed=9):
    rng = np.random.default_rng(seed)
    hist = rng.uniform(-1, 1, size=(B, T + 1, A, N, 5)).astype(np.float32)
    hist[..., 0] = 1.0
    hist[:, :, :, 20:] = 0.0
    term = np.zeros((B, T + 1, A, 1), dtype=np.uint8)
    term[1, 4:, 0] = 1
    term[B - 2, 2:, 1] = 1
    return dict(history=hist, attention_latent=rng.uniform(-1, 1, size=(B, T + 1, A, N, 32)).astype(np.float32),
                behavior_latent=rng.dirichlet(np.ones(8), size=(B, T + 1, A, N)).astype(np.float32),
                rnn_states_actors=rng.uniform(-1, 1, size=(B, T + 1, A, 64)).astype(np.float32),
                rnn_states_critics=rng.uniform(-1, 1, size=(B, T + 1, A, 64)).astype(np.float32),
                actions=rng.integers(0, 5, size=(B, T + 1, A, 1)), avail_actions=np.ones((B, T + 1, A, 5), dtype=np.int64),
                reward=(rng.normal(size=(B, T + 1, A, 1)) * 2).astype(np.float32), terminated=term)


def _worker(rank, world, port, out):
    import torch.distributed as dist
    from iplan_b200.config import make_args
    from tests.test_gpu_learner import build
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        Bg, T, A, N = 16, 7, 5, 55
        data = _data(Bg, T, A, N)
        torch.manual_seed(4)
        from iplan_b200.modules.flat import ParamStack
        F = N * 45 + 10
        a0, c0 = ParamStack("actor", A, (F, 5)), ParamStack("critic", A, (F,))
        with torch.no_grad():
            for n in a0.nets:
                n.act.action_out.linear.weight.mul_(30.0)
        actors = [{k: v.clone() for k, v in n.state_dict().items()} for n in a0.nets]
        critics = [{k: v.clone() for k, v in n.state_dict().items()} for n in c0.nets]
        Bl = Bg // world
        common = dict(episode_limit=T, ppo_epoch=3, use_cuda=True, device="cuda", batch_size=Bg - 1)
        # sharded learner: this rank's episodes only, global batch_size
        a_sh = make_args("highway", buffer_size=Bl, batch_size_run=Bl, **common)
        shard = {k: v[rank * Bl:(rank + 1) * Bl] for k, v in data.items()}
        batch, mac, learner, _ = build(a_sh, shard, actors, critics)
        learner.insert_episode_batch(batch)
        learner.train(0)
        torch.cuda.synchronize()
        res = None
        if rank == 0:
            a_full = make_args("highway", buffer_size=Bg, batch_size_run=Bg, **common)
            batch2, mac2, learner2, _ = build(a_full, data, actors, critics)
            learner2.use_dist = False
            learner2.insert_episode_batch(batch2)
            learner2.train(0)
            torch.cuda.synchronize()
            da = float((mac.actor_stack.flat - mac2.actor_stack.flat).abs().max())
            dc = float((mac.critic_stack.flat - mac2.critic_stack.flat).abs().max())
            moved = float((mac2.actor_stack.flat.cpu() - a0.flat).abs().max())
            res = dict(da=da, dc=dc, moved=moved, info=learner.train_info, info2=learner2.train_info)
        # both ranks must hold identical weights after the update
        w = mac.actor_stack.flat.clone()
        dist.broadcast(w, 0)
        same = float((w - mac.actor_stack.flat).abs().max())
        if rank == 0:
            res["rank_spread"] = same
            out.put(res)
        else:
            out.put(dict(rank_spread=same))
    finally:
        dist.destroy_process_group()


def test_two_gpu_sharded_update_equals_single_gpu():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 CUDA devices")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    r0 = [r for r in res if "da" in r][0]
    print("[multi-gpu]", {k: v for k, v in r0.items() if k not in ("info", "info2")})
    assert all(r["rank_spread"] == 0.0 for r in res)            # replicas stay bit-identical
    assert r0["moved"] > 1e-4                                    # the update did something
    assert r0["da"] < 2e-5 and r0["dc"] < 2e-5                   # sharded == single-rank
    for k in ("value_loss", "policy_loss", "dist_entropy", "ratio", "actor_grad_norm", "critic_grad_norm"):
        assert abs(r0["info"][k] - r0["info2"][k]) < 1e-4 * max(1.0, abs(r0["info2"][k])), k
