"""The restated YAML merge (iplan_b200/config.py) equals the reference's merged
config captured in tests/golden/config.pt (made from /root/reference/config/*)."""
import os

import torch

from iplan_b200.config import controller_input_dim, make_args, merged_config


def test_merged_config_matches_reference(golden_dir):
    ref = torch.load(os.path.join(golden_dir, "config.pt"), weights_only=False)
    for env in ("highway", "MPE"):
        mine = merged_config(env)
        for k, v in ref[env].items():
            assert mine[k] == v, (env, k)


def test_input_dims():
    assert controller_input_dim(make_args("highway")) == 2485
    assert controller_input_dim(make_args("MPE")) == 272
    a = make_args("highway")
    assert (a.max_vehicle_num, a.n_agents, a.episode_limit) == (55, 5, 90)
