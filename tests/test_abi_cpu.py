"""CPU-side checks of the C-ABI library: it loads without a GPU, exports every symbol
include/iplan_b200.h declares, and its parameter layouts agree with the host-side
module specs.  No compute calls."""
import ctypes
import os
import re

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from iplan_b200 import _lib
    header = open(os.path.join(ROOT, "include", "iplan_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    names = set(re.findall(r"\b(iplan_[a-z0-9_]+)\s*\(", header))
    assert len(names) >= 10
    for n in sorted(names):
        assert hasattr(_lib.lib, n), f"{n} declared in include/iplan_b200.h but not exported"
        assert n in _lib._SIGNATURES, f"{n} has no ctypes signature in iplan_b200/_lib.py"
    assert _lib.ABI_VERSION == int(re.search(r"#define IPLAN_ABI_VERSION (\d+)", header).group(1))


def test_pipeline_pieces_are_whole_waves_with_a_short_last_piece():
    """_lib.wave_chunks: the env pieces of the native GAT_latent_update pipeline cover [0, n_envs), are whole waves of a
    one-CTA-per-SM kernel (their waves add up to those of a single launch) and end with the shortest piece."""
    from iplan_b200 import _lib
    _lib._sm_count = 148
    try:
        for n_envs in (1, 59, 60, 128, 130, 256, 512, 513, 1024, 4096):
            for agents in (1, 3, 5, 8, 16):
                ctas = lambda e: ((e + 1) // 2) * agents                      # K1's grid
                waves = lambda e: -(-ctas(e) // 148) if e else 0
                ends = _lib.wave_chunks(n_envs, ctas)
                assert ends[-1] == n_envs and all(b > a for a, b in zip([0] + ends, ends)) and 1 <= len(ends) <= 3
                sizes = [b - a for a, b in zip([0] + ends, ends)]
                assert sum(waves(x) for x in sizes) <= waves(n_envs), (n_envs, agents, ends)
                if len(ends) > 1:
                    assert waves(sizes[-1]) == 1 and sizes[-1] == min(sizes), (n_envs, agents, ends)
        assert _lib.wave_chunks(512, lambda e: ((e + 1) // 2) * 5) == [236, 472, 512]
    finally:
        _lib._sm_count = None


def test_layouts_match_module_specs():
    from iplan_b200.modules.flat import ParamStack
    for kind, dims in (("gat", (13,)), ("beh", (5, 8)), ("actor", (2485, 5)), ("critic", (2485,)),
                       ("actor", (272, 5)), ("gat", (12,))):
        s = ParamStack(kind, 2, dims)
        end = 0
        for (name, shape), off in zip(s.spec, s.offsets):
            n = 1
            for d in shape:
                n *= d
            assert off % 4 == 0 and off >= end, (kind, name)
            end = off + n
        assert end <= s.total and s.total % 4 == 0
        sd = s.nets[1].state_dict()
        assert [k for k in sd] == [n for n, _ in s.spec]
    # parameter counts of the reference modules
    a = ParamStack("actor", 1, (2485, 5))
    assert sum(p.numel() for p in a.nets[0].parameters()) == 198191
    c = ParamStack("critic", 1, (2485,))
    assert sum(p.numel() for p in c.nets[0].parameters()) == 197935      # incl. 4 frozen PopArt scalars
    g = ParamStack("gat", 1, (13,))
    assert sum(p.numel() for p in g.nets[0].parameters()) == 28834


def test_state_dict_roundtrip_with_reference_checkpoint_format(tmp_path, golden_dir):
    from iplan_b200.modules.flat import ParamStack
    g = torch.load(os.path.join(golden_dir, "rollout_modules.pt"), weights_only=False)["mpe"]
    s = ParamStack("critic", 3, (272,))
    s.nets[2].load_state_dict(g["critics"][2])
    torch.save(s.nets[2].state_dict(), tmp_path / "critic_2.th")
    back = torch.load(tmp_path / "critic_2.th", weights_only=False)
    for k, v in g["critics"][2].items():
        assert torch.equal(back[k], v), k


def test_episode_batch_cpu_semantics():
    """Non-packed (CPU) EpisodeBatch keeps the reference's update/view_as/filled semantics."""
    from iplan_b200.components.episode_buffer import EpisodeBatch
    from iplan_b200.components.transforms import OneHot
    scheme = {"actions": {"vshape": (1,), "group": "agents", "dtype": torch.long},
              "rnn": {"vshape": (4,), "group": "agents"}, "reward": {"vshape": (1,), "group": "agents"}}
    b = EpisodeBatch(scheme, {"agents": 2}, 3, 5, preprocess={"actions": ("actions_onehot", [OneHot(3)])})
    b.update({"rnn": torch.arange(24.).view(1, 3, 2, 4)}, ts=1)         # [1,B,A,R] -> [B,1,A,R]
    assert torch.equal(b["rnn"][:, 1], torch.arange(24.).view(3, 2, 4))
    assert b["filled"][:, 1].sum() == 3 and b["filled"].sum() == 3 and int(b.max_t_filled()) == 1
    b.update({"actions": [[1, 2], [0, 0], [2, 1]]}, ts=0, mark_filled=False)
    assert b["actions_onehot"][0, 0].tolist() == [[0, 1, 0], [0, 0, 1]]
    assert b["filled"].sum() == 3
    sub = b[1:3, 0:2]
    assert sub.batch_size == 2 and sub.max_seq_length == 2 and sub["rnn"].shape == (2, 2, 2, 4)
    try:
        b.update({"nope": 1})
        assert False
    except KeyError:
        pass


def test_prediction_optimizer_state_dict_is_torch_adam_compatible():
    """``pred_optimizer_{i}_opt.th`` (reference nova/prediction_policy.py:262): the flat Adam moments of the GAT + decoder
    stacks export / import in torch.optim.Adam's state_dict format (parameter ids: GAT tensors, then decoder tensors)."""
    from types import SimpleNamespace
    from iplan_b200.modules.flat import ParamStack
    from iplan_b200.nova.prediction_policy import _PredAdam
    gat, dec = ParamStack("gat", 2, (13,)), ParamStack("pdec", 2, (5,))
    z = lambda t: torch.zeros_like(t)
    state = dict(m_gat=torch.randn_like(gat.flat), v_gat=torch.rand_like(gat.flat), m_dec=torch.randn_like(dec.flat),
                 v_dec=torch.rand_like(dec.flat), step=3)
    owner = SimpleNamespace(stack=gat, dec_stack=dec, device=torch.device("cpu"), _learn=state, _learn_state=lambda: state,
                            args=SimpleNamespace(lr_predict=2e-5, optim_eps=1e-5, weight_decay=0))
    sd = _PredAdam(owner, 1).state_dict()
    params = [torch.nn.Parameter(torch.zeros(shape)) for _, shape in gat.spec + dec.spec]
    opt = torch.optim.Adam(params, lr=2e-5, eps=1e-5)
    opt.load_state_dict(sd)                                    # torch accepts the format
    assert len(sd["state"]) == len(params) == 28
    off, shape = gat.named_offsets()["q.weight"]
    pid = [n for n, _ in gat.spec].index("q.weight")
    assert torch.equal(opt.state[params[pid]]["exp_avg"], state["m_gat"][1, off:off + 1024].view(shape))
    # round trip into a fresh owner
    state2 = dict(m_gat=z(gat.flat), v_gat=z(gat.flat), m_dec=z(dec.flat), v_dec=z(dec.flat), step=0)
    owner2 = SimpleNamespace(stack=gat, dec_stack=dec, device=torch.device("cpu"), _learn=state2, _learn_state=lambda: state2, args=owner.args)
    _PredAdam(owner2, 1).load_state_dict(opt.state_dict())
    assert state2["step"] == 3
    for (name, shape), off in zip(dec.spec, dec.offsets):
        n = int(torch.tensor(shape).prod()) if len(shape) else 1
        assert torch.equal(state2["v_dec"][1, off:off + n], state["v_dec"][1, off:off + n]), name
