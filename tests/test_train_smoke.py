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


@pytest.mark.parametrize("use_graph", [False, True])
def test_boxworld_rollout_in_env_ranges_equals_whole_batch_rollout(use_graph):
    """The rollout of env ranges on concurrent streams (``rollout_chains``) must produce what the single-chain rollout
    produces: per-env random streams are keyed by the env index, the level-seed counter is consumed in env order
    (events between the ranges), and the reward normalisation runs once per rollout over all envs."""
    from tpp_b200.boxworld.box_world_env_vec import create_bw_env
    hp = dict(n_envs=512, grid_size=6, goal_length=2, num_distractor=1, distractor_length=1, max_steps=12)
    outs = []
    for chains in (1, 4, 2):
        env = create_bw_env(None, hp)
        agent = _agent(env, (3, 8, 8), 192, 24, 512, use_graph, rollout_chains=chains)
        assert len(agent._env_ranges(env, agent.storage)) == chains
        agent.train(24 * 512 * 3)
        st = agent.storage
        outs.append(dict(frames=st.frames.clone(), act=st.act_i32.clone(), done=st.done_u8.clone(),
                         raw=st.env_rew_i32.clone(), rew=st.rew.clone(), value=st.value.clone(), logp=st.logp.clone(),
                         seed=env.env._seed_counter.clone(), rms=env._rms.clone(), flat=agent.policy.flat.clone()))
    assert (outs[0]["done"].sum() > 50) and (outs[0]["raw"] != 0).any()       # resets and rewards do occur
    for o in outs[1:]:
        for k in ("frames", "act", "done", "raw", "seed"):
            assert torch.equal(outs[0][k], o[k]), k
        for k in ("rew", "value", "logp", "rms", "flat"):
            torch.testing.assert_close(outs[0][k], o[k], rtol=1e-5, atol=1e-6, msg=k)


@pytest.mark.parametrize("parallel", [False, True])                 # one CTA walking the steps / four parallel launches
@pytest.mark.parametrize("N,ld", [(1000, 1024), (5000, 5024)])      # register path (<= 4096 envs) and the generic one
def test_vecnormalize_rollout_kernel_equals_per_step_kernel(N, ld, parallel):
    from tpp_b200 import _lib
    T = 37
    scratch = torch.zeros(T * ld + 3 * T if parallel else 1, dtype=torch.float64, device="cuda")
    g = torch.Generator().manual_seed(3)
    raw = torch.randint(-1, 12, (T, ld), generator=g, dtype=torch.int32).cuda()
    done = (torch.rand(T, ld, generator=g) < 0.05).to(torch.uint8).cuda()
    ret_a = torch.zeros(N, dtype=torch.float64, device="cuda"); ret_b = ret_a.clone()
    rms_a = torch.tensor([0.0, 1.0, 1e-4], dtype=torch.float64, device="cuda"); rms_b = rms_a.clone()
    out_a = torch.zeros(T, ld, device="cuda"); out_b = out_a.clone(); raw_f = out_a.clone()
    for t in range(T):
        _lib.call("tpp_vecnormalize_step", _lib.ptr(ret_a), _lib.ptr(rms_a), _lib.ptr(raw[t]), 1, _lib.ptr(done[t]),
                  _lib.ptr(out_a[t]), N, 0.999, 10.0, 1e-8, _lib.stream_ptr())
    _lib.call("tpp_vecnormalize_rollout", _lib.ptr(ret_b), _lib.ptr(rms_b), _lib.ptr(raw), _lib.ptr(done),
              _lib.ptr(out_b), _lib.ptr(raw_f), T, N, ld, 0.999, 10.0, 1e-8, _lib.ptr(scratch) if parallel else None,
              scratch.numel() if parallel else 0, _lib.stream_ptr())
    torch.testing.assert_close(out_a[:, :N], out_b[:, :N], rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(ret_a, ret_b, rtol=1e-12, atol=1e-12)
    torch.testing.assert_close(rms_a, rms_b, rtol=1e-12, atol=0)
    assert torch.equal(raw_f[:, :N], raw[:, :N].float())
