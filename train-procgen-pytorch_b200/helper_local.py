"""The hot-path slice of the reference's ``helper_local.py`` / ``train.py`` wiring, so that the YAML sets of
``hyperparams/procgen/config.yml`` are consumed UNCHANGED (SURVEY 5.6, 8b "Config"):

* ``get_hyperparams(param_name)``                   -- helper_local.py:207-210
* ``initialize_model(device, env, hyperparameters)`` -- helper_local.py:213-330 ('impala' and 'mlpmodel' architectures)
* ``initialize_storage(...)``                        -- helper_local.py:1123-1141
* ``train_ppo(args, hyperparameters)``               -- train.py:125-268 (env -> logger -> model -> storage -> PPO ->
                                                        optional checkpoint load -> ``agent.train``)

Same key names and defaults; the whole set is splatted into the env constructors (signature-filtered,
discrete_env/pre_vec_env.py:209-219) and into ``PPO(**hyperparameters)``, unknown keys ignored, exactly like upstream.
The YAML file itself is the user's: ``config_path`` / ``$TPP_CONFIG_YML`` / ``./hyperparams/procgen/config.yml``.
"""
from __future__ import annotations

import os

import numpy as np
import torch
import yaml

from .agents.ppo import PPO
from .common.env.env_constructor import get_env_constructor
from .common.logger import Logger
from .common.model import ImpalaModel, MLPModel
from .common.policy import CategoricalPolicy
from .common.storage import Storage


def config_path(path=None):
    for p in (path, os.environ.get("TPP_CONFIG_YML"), os.path.join(os.getcwd(), "hyperparams/procgen/config.yml")):
        if p and os.path.exists(p):
            return p
    raise FileNotFoundError("hyperparams/procgen/config.yml not found: pass config_path= or set TPP_CONFIG_YML to the "
                            "reference checkout's file")


def get_hyperparams(param_name, path=None):
    with open(config_path(path), "r") as f:
        return yaml.safe_load(f)[param_name]


def initialize_model(device, env, hyperparameters, in_channels=None):
    """-> (model, observation_shape, policy); the policy lives on ``device`` and is flattened for the engine."""
    observation_shape = tuple(env.observation_space.shape)
    architecture = hyperparameters.get("architecture", "impala")
    if in_channels is None:
        in_channels = observation_shape[0]
    action_size = env.action_space.n
    if architecture == "impala":
        model = ImpalaModel(in_channels=in_channels, output_dim=hyperparameters.get("output_dim", 256),
                            latent_dim=hyperparameters.get("latent_dim", 32), input_hw=observation_shape[1:])
    elif architecture == "mlpmodel":
        model = MLPModel(in_channels, hyperparameters.get("depth", 4), hyperparameters.get("mid_weight", 64),
                         hyperparameters.get("latent_size", 256))
    else:
        raise NotImplementedError(f"architecture '{architecture}' is outside the hot path (SURVEY 8a10: impala, mlpmodel)")
    recurrent = hyperparameters.get("recurrent", False)
    policy = CategoricalPolicy(model, recurrent, action_size)
    policy.to(device)
    policy.flatten_(device)
    policy.device = device
    return model, observation_shape, policy


def initialize_storage(args, device, hidden_state_dim, n_envs, n_steps, observation_shape):
    storage = Storage(observation_shape, hidden_state_dim, n_steps, n_envs, device)
    storage_valid = Storage(observation_shape, hidden_state_dim, n_steps, n_envs, device) \
        if getattr(args, "use_valid_env", False) else None
    return storage, storage_valid, None


class Args:
    """Stand-in for the argparse namespace of train.py (only the fields this path reads)."""

    def __init__(self, **kw):
        self.seed, self.num_levels, self.use_valid_env, self.model_file = 6033, 500, False, None
        self.num_timesteps, self.num_checkpoints, self.device, self.logdir = None, 0, "cuda", None
        self.__dict__.update(kw)


def train_ppo(args, hyperparameters, env_name, train=True, **agent_kwargs):
    """train.py:125-268 for algo 'ppo'.  Returns the agent (after ``agent.train(args.num_timesteps)`` when asked)."""
    device = torch.device(getattr(args, "device", "cuda") if getattr(args, "device", "cuda") != "gpu" else "cuda")
    hyperparameters = dict(hyperparameters)
    n_steps, n_envs = hyperparameters.get("n_steps", 256), hyperparameters.get("n_envs", 256)
    max_steps = hyperparameters.get("max_steps", 10 ** 3)
    if hyperparameters.get("algo", "ppo") != "ppo":
        raise NotImplementedError("only algo 'ppo' is on the hot path")
    create_venv = get_env_constructor(env_name)
    env_hp = dict(hyperparameters, device=str(device))
    env = create_venv(args, env_hp)
    env_valid = create_venv(args, env_hp, is_valid=True) if getattr(args, "use_valid_env", False) else None
    logdir = getattr(args, "logdir", None)
    if logdir:
        os.makedirs(logdir, exist_ok=True)
        np.save(os.path.join(logdir, "hyperparameters.npy"), hyperparameters)
    model, observation_shape, policy = initialize_model(device, env, hyperparameters)
    logger = Logger(n_envs, logdir)
    logger.max_steps = max_steps
    storage, storage_valid, _ = initialize_storage(args, device, model.output_dim, n_envs, n_steps, observation_shape)
    agent = PPO(env, policy, logger, storage, device, getattr(args, "num_checkpoints", 0), env_valid=env_valid,
                storage_valid=storage_valid, **hyperparameters, **agent_kwargs)
    if getattr(args, "model_file", None) is not None:            # train.py:257-263
        checkpoint = torch.load(args.model_file, map_location=device)
        agent.policy.load_state_dict(checkpoint["model_state_dict"])
        agent.optimizer.load_state_dict(checkpoint["optimizer_state_dict"])
    if train and getattr(args, "num_timesteps", None):
        agent.train(args.num_timesteps)
    return agent
