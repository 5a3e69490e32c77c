"""Run configuration for the iPLAN hot path (host side, plain Python).

The reference merges three YAML files with a "first key wins" rule
(/root/reference/main.py:59-69, :83-100: config/default.yaml, then
config/envs/<env>.yaml, then config/algs/ippo.yaml) and turns the result into a
``SimpleNamespace`` (/root/reference/run_ippo.py:39); ``run_sequential`` then
derives n_agents / max_vehicle_num / shapes from the runner
(/root/reference/run_ippo.py:136-147).  The YAML files do not travel to the GPU
box, so the merged values are restated here; ``tests/golden/make_golden.py``
re-derives them from the reference's YAMLs and ``tests/test_config.py`` pins this
table against the committed result.
"""
from types import SimpleNamespace

# config/default.yaml (first wins) ------------------------------------------------
_DEFAULT = dict(
    runner="parallel", mac="dcntrl", env="MPE", difficulty="hard", env_args={},
    batch_size_run=8, test_nepisode=20, test_interval=20000, test_greedy=True,
    log_interval=20000, runner_log_interval=20000, learner_log_interval=20000,
    t_max=2000000, use_cuda=True, buffer_cpu_only=True, checkpoint_paths=[""],
    num_test_episodes=8, use_tensorboard=True, save_model=True,
    save_model_interval=100000, checkpoint_path="", evaluate=False, load_step=0,
    save_replay=False, local_results_path="results",
    gamma=0.99, batch_size=255, buffer_size=256, lr=0.0005, critic_lr=0.0005,
    optim_alpha=0.99, optim_eps=0.00001, grad_norm_clip=10,
    agent="ippo", critic="ippo", rnn_hidden_dim=64, mlp_hidden_dim=64,
    obs_agent_id=True, obs_last_action=True, repeat_id=1, label="",
    log_prefix="ippo_GAT_behavior_stable_H_",
    max_history_len=10,
    Behavior_enable=True, Behavior_warmup=20000, encoder_rnn_dim=32,
    num_encoder_layer=1, latent_dim=8, decoder_rnn_dim=64, num_decoder_layer=1,
    lr_behavior=0.0001, decoder_dropout=0.1, soft_update_enable=True,
    soft_update_coef=0.1, behavior_variation_penalty=0, thres_small_variation=0.005,
    behavior_fully_connected=False,
    GAT_enable=True, GAT_use_behavior=True, GAT_warmup=20000, GAT_hidden_dim=32,
    attention_dim=32, teacher_forcing_ratio=0, pred_batch_size=64,
    lr_predict=0.00002, pred_dropout=0.1, pred_length=5,
    use_max_grad_norm=True, max_grad_norm=10.0,
    animation_enable=False, metrics_enable=False,
)

# config/envs/highway.yaml ---------------------------------------------------------
_ENV_HIGHWAY = dict(
    env="highway", n_lane=8, n_actions=5, obs_shape_single=5, n_agents=5,
    n_other_vehicles=50, n_obs_vehicles=15, episode_limit=90, scaling=2.0,
    screen_height=300, screen_width=1200, max_history_len=10,
)

# config/envs/simple_spread_Hetero.yaml ("easy": 3 agents, 3 landmarks) --------------
_ENV_MPE = dict(
    env="MPE", scenario_name="simple_spread_Hetero", num_agents=3, num_landmarks=3,
    episode_length=50, num_normal_agents=1, num_tiny_agents=1, num_bulky_agents=1,
    num_random_agents=0, world_size=1.0, obs_shape_single=4, init_sample_size=5,
    n_actions=5,
)

# config/algs/ippo.yaml ------------------------------------------------------------
_ALG_IPPO = dict(
    weight_decay=0, ppo_epoch=15, use_clipped_value_loss=True,
    use_linear_lr_decay=False, clip_param=0.2, num_mini_batch=1,
    data_chunk_length=10, value_loss_coef=0.5, entropy_coef=0.01,
    use_max_grad_norm=True, max_grad_norm=10.0, use_gae=True, gae_lambda=0.95,
    use_proper_time_limits=True, use_huber_loss=True, huber_delta=10.0,
    gain=0.01, use_orthogonal=True, use_policy_active_masks=True,
    use_value_active_masks=True, use_recurrent_policy=True, recurrent_N=1,
    rnn_hidden_dim=64, use_ReLU=True, stacked_frames=1, layer_N=1,
    mlp_hidden_dim=64, use_feature_normalization=True, use_popart=True,
    action_selector="epsilon_greedy", epsilon_start=1.0, epsilon_finish=0.05,
    epsilon_anneal_time=50000, agent_output_type="None",  # YAML "None" is a string
    runner="ippo", learner="ippo_learner", name="ippo",
)


def merged_config(env="highway"):
    """default.yaml, then the env file, then the alg file; a key already present is
    kept (main.py:59-69 ``recursive_dict_update`` only fills missing keys)."""
    cfg = dict(_DEFAULT)
    for layer in ((_ENV_HIGHWAY if env == "highway" else _ENV_MPE), _ALG_IPPO):
        for k, v in layer.items():
            # NB the env file is picked by default.yaml's own ``env`` key, which the
            # user edits to "highway"/"MPE" (main.py:85-92); here ``env`` is the
            # selector, so it is taken from the chosen env file.  Every other
            # duplicated key keeps default.yaml's value.
            if k not in cfg or k == "env":
                cfg[k] = v
    return cfg


def make_args(env="highway", **overrides):
    """``args`` namespace as ``run_ippo.run`` + ``run_sequential`` would hold it
    after the runner reported its env_info (run_ippo.py:136-147)."""
    cfg = merged_config(env)
    cfg.update(overrides)
    a = SimpleNamespace(**cfg)
    if a.env == "highway":
        a.max_vehicle_num = a.n_other_vehicles + a.n_agents            # run_ippo.py:145
        a.state_shape = a.obs_shape_single * a.max_vehicle_num         # runner.get_env_info
        a.obs_shape = a.obs_shape_single * a.n_obs_vehicles
    else:
        a.n_agents = a.num_agents
        a.max_vehicle_num = a.num_landmarks + a.n_agents + a.num_random_agents  # :147
        a.episode_limit = a.episode_length
        a.state_shape = a.obs_shape_single * a.max_vehicle_num
        a.obs_shape = a.obs_shape_single * a.max_vehicle_num
    for k, v in overrides.items():       # explicit overrides of derived values win
        setattr(a, k, v)
    if not hasattr(a, "device"):
        a.device = "cuda" if a.use_cuda else "cpu"
    return a


def controller_input_dim(args):
    """DcntrlMAC._get_input_shape (controllers/dcntrl_controller.py:215-232) for the
    full iPLAN setting (GAT + behaviour enabled, last action + agent id appended)."""
    n = args.max_vehicle_num
    return n * (args.obs_shape_single + args.attention_dim + args.latent_dim) \
        + args.n_actions + args.n_agents
