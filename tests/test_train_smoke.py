"""GPU: end-to-end PPO.train on the device (fused rollout, CUDA-graph replay, GAE, update) for small shapes."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _agent(env, obs_shape, in_dim, n_steps, n_envs, use_graph, **kw):
    from tpp_b200.agents.ppo import PPO
    from tpp_b200.common.model import MLPModel
    from tpp_b200.common.policy import CategoricalPolicy
    from tpp_b200.common.storage import Storage
    torch.manual_seed(6033)
    pol = CategoricalPolicy(MLPModel(in_dim, 4, 64, 32), False, env.action_space.n).to("cuda").flatten_()
    st = Storage(obs_shape, 32, n_steps, n_envs, "cuda")
    return PPO(env, pol, None, st, "cuda", 0, n_steps=n_steps, n_envs=n_envs, epoch=2, n_minibatch=4,
               mini_batch_size=8192, learning_rate=5e-4, entropy_coef=0.02, use_cuda_graph=use_graph, **kw)


@pytest.mark.parametrize("use_graph", [False, True])
def test_cartpole_train_runs_and_learns_something(use_graph):
    from tpp_b200.discrete_env.cartpole_pre_vec import CartPoleVecEnv
    env = CartPoleVecEnv(n_envs=256, seed=6033)
    agent = _agent(env, (9,), 9, 64, 256, use_graph)
    w0 = agent.policy.flat.clone()
    agent.train(64 * 256 * 6)
    assert agent.t == 64 * 256 * 6
    assert torch.isfinite(agent.policy.flat).all() and not torch.equal(agent.policy.flat, w0)
    rew, done, _ = agent.storage.fetch_log_data()
    assert rew.shape == (64, 256) and (rew == 1).all() and 0 < done.mean() < 0.2
    # rollout bookkeeping: log-probs are log-probs, values finite, obs slots chained (slot 0 == previous slot T)
    assert (agent.storage.log_prob_act_batch <= 0).all() and torch.isfinite(agent.storage.value_batch).all()
    assert torch.equal(agent.storage.obs_slot(0), agent.storage.obs_slot(64))
    assert agent.optimizer.step_count == 6 * 2 * 4


def test_graph_and_eager_rollouts_agree():
    """Same seeds -> the CUDA-graph replay produces exactly the rollout the eager launch sequence produces."""
    from tpp_b200.discrete_env.acrobot_pre_vec import AcrobotVecEnv
    outs = []
    for use_graph in (False, True):
        env = AcrobotVecEnv(n_envs=128, seed=1, max_steps=20)
        agent = _agent(env, (14,), 14, 32, 128, use_graph)
        agent.train(32 * 128 * 4)
        outs.append((agent.policy.flat.clone(), agent.storage.obs_fm.clone(), agent.storage.act_i32.clone()))
    assert torch.equal(outs[0][2], outs[1][2]) and torch.equal(outs[0][1], outs[1][1])
    np.testing.assert_allclose(outs[0][0].cpu().numpy(), outs[1][0].cpu().numpy(), rtol=1e-4, atol=1e-6)


def test_boxworld_train_runs():
    from tpp_b200.boxworld.box_world_env_vec import create_bw_env
    hp = dict(n_envs=64, grid_size=6, goal_length=2, num_distractor=1, distractor_length=1, max_steps=50)
    env = create_bw_env(None, hp)
    agent = _agent(env, (3, 8, 8), 192, 32, 64, True)
    agent.train(32 * 64 * 3)
    assert torch.isfinite(agent.policy.flat).all()
    rew, done, _ = agent.storage.fetch_log_data()
    assert set(np.unique(rew)).issubset({-1.0, 0.0, 1.0, 10.0, 11.0})
