"""GPU: IMPALA-CNN policy (reference common/model.py:134-208) through the hand-written engine (im2col + tcgen05 GEMM,
matmul="tf32x3") and through the tests' cross-check engine (cuDNN via torch autograd, tests/torch_engine.py), both with this
repo's gather / fused loss / clip+Adam kernels, and the host-stepped env path
(Procgen-style numpy VecEnv staged through Storage.store).  Parity: one optimize() against the torch-CPU oracle."""
import numpy as np
import pytest
import torch

from oracle import ppo as oppo

pytestmark = pytest.mark.gpu


def _impala_agent(T, N, A, hw=(64, 64), **kw):
    from tpp_b200.agents.ppo import PPO
    from tpp_b200.common.model import ImpalaModel
    from tpp_b200.common.policy import CategoricalPolicy
    from tpp_b200.common.storage import Storage
    torch.manual_seed(3)
    pol = CategoricalPolicy(ImpalaModel(3, input_hw=hw), False, A).to("cuda").flatten_()
    st = Storage((3, *hw), 256, T, N, "cuda")
    if kw.get("matmul") == "library":
        from torch_engine import TorchModuleEngine
        kw.pop("matmul")
        kw["engine"] = TorchModuleEngine(pol, A, (3, *hw))
    agent = PPO(kw.pop("env", None), pol, None, st, "cuda", 0, n_steps=T, n_envs=N, epoch=1, n_minibatch=2,
                mini_batch_size=64, learning_rate=5e-4, entropy_coef=0.01, **kw)
    return agent, pol, st


@pytest.mark.parametrize("matmul,fs_coef", [("tf32x3", 0.0), ("library", 0.0), ("tf32x3", 0.3), ("library", 0.3)])
def test_impala_optimize_matches_oracle(matmul, fs_coef):
    """fs_coef != 0: the feature-sparsity term of the loss (common/model.py:203-208, agents/ppo.py:164-169) and its
    gradient on the hand-written engine (tpp_feature_sparsity / _grad)."""
    from tpp_b200.common.engine import ImpalaEngineTC
    from torch_engine import TorchModuleEngine
    T, N, A = 8, 16, 15
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    agent, pol, st = _impala_agent(T, N, A, matmul=matmul, fs_coef=fs_coef)
    assert isinstance(agent.engine, ImpalaEngineTC if matmul == "tf32x3" else TorchModuleEngine)
    g = torch.Generator().manual_seed(0)
    frames = torch.randint(0, 256, (T + 1, N, 64, 64, 3), generator=g, dtype=torch.uint8)
    st.frames.copy_(frames.cuda())
    st.act_i32[:, :N] = torch.randint(0, A, (T, N), generator=g).cuda().int()
    st.logp[:, :N] = (-torch.rand(T, N, generator=g) * 2 - 0.5).cuda()
    st.value[:, :N] = (torch.randn(T + 1, N, generator=g) * 0.3).cuda()
    st.rew[:, :N] = torch.randn(T, N, generator=g).cuda()
    st.done_u8[:, :N] = (torch.rand(T, N, generator=g) < 0.1).cuda().to(torch.uint8)
    st.compute_estimates(0.999, 0.95, True, True)
    # oracle: same weights, same data, same torch seed for the minibatch permutation
    ref = oppo.OraclePolicy(oppo.OracleImpala(3), A)
    ref.load_state_dict({k.replace("embedder.", "embedder."): v.detach().cpu() for k, v in pol.state_dict().items()})
    data = dict(obs=(frames[:-1].permute(0, 1, 4, 2, 3).float() / 255.0).reshape(T * N, 3, 64, 64),
                act=st.act_batch.cpu().reshape(-1), old_logp=st.log_prob_act_batch.cpu().reshape(-1),
                old_value=st.value_batch[:-1].cpu().reshape(-1), ret=st.return_batch.cpu().reshape(-1),
                adv=st.adv_batch.cpu().reshape(-1))
    opt = oppo.make_adam(ref, 5e-4)
    torch.manual_seed(99)
    logs = oppo.optimize(ref, opt, data, T, N, epoch=1, n_minibatch=2, mini_batch_size=64, grad_clip_norm=0.5,
                         eps_clip=0.2, value_coef=0.5, entropy_coef=0.01, fs_coef=fs_coef)
    torch.manual_seed(99)
    summary = agent.optimize()
    np.testing.assert_allclose(summary["Loss/total"], np.mean([l["total"] for l in logs]), rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(summary["Loss/entropy"], np.mean([l["entropy"] for l in logs]), rtol=2e-4)
    # IMPALA reports the feature sparsity every minibatch (common/model.py:203-208)
    np.testing.assert_allclose(summary["Loss/feature_sparsity"], np.mean([l["fs"] for l in logs]), rtol=1e-4)
    for (k, p), (_, q) in zip(pol.state_dict().items(), ref.state_dict().items()):
        np.testing.assert_allclose(p.cpu().numpy(), q.numpy(), rtol=5e-3, atol=2e-5, err_msg=k)


class FakeProcgen:
    """Host-stepped numpy VecEnv with Procgen's observation contract (float NCHW in [0,1] after the wrappers)."""

    def __init__(self, n, hw=(64, 64)):
        from tpp_b200.discrete_env.pre_vec_env import Box, Discrete
        self.n, self.hw = n, hw
        self.observation_space = Box(np.zeros((3, *hw)), np.ones((3, *hw)))
        self.action_space = Discrete(15)
        self.rng = np.random.default_rng(0)

    def _obs(self):
        return self.rng.integers(0, 256, (self.n, 3, *self.hw)).astype(np.float32) / 255.0

    def reset(self):
        return self._obs()

    def step(self, act):
        assert act.shape == (self.n,) and act.dtype == np.int64
        rew = (self.rng.random(self.n) < 0.01) * 10.0
        done = self.rng.random(self.n) < 0.01
        return self._obs(), rew, done, [{"env_reward": r} for r in rew]

    def close(self):
        pass


def test_host_env_staging_loop_trains():
    """Config C4 in miniature: host env -> pinned staging into the uint8 GPU rollout -> GPU GAE + update."""
    env = FakeProcgen(16)
    agent, pol, st = _impala_agent(8, 16, 15, env=env)
    w0 = pol.flat.clone()
    agent.train(8 * 16 * 2)
    assert agent.t == 8 * 16 * 2 and torch.isfinite(pol.flat).all() and not torch.equal(pol.flat, w0)
    assert st.frames.dtype == torch.uint8 and int(st.frames.max()) > 200


@pytest.mark.parametrize("matmul", ["tf32x3", "fp32"])
def test_mlp_on_frames_optimize_matches_oracle(matmul):
    """MLP policy on uint8 frames (the Box-World bench shape).  With the tensor-core engine the first layer reads the
    integer pixel values (exact TF32 operand, 1/255 folded into its weight copy and its weight-gradient alpha): one
    optimize() must still match the torch-CPU oracle that sees float frames / 255."""
    from tpp_b200.agents.ppo import PPO
    from tpp_b200.common.engine import MLPEngineTC
    from tpp_b200.common.model import MLPModel
    from tpp_b200.common.policy import CategoricalPolicy
    from tpp_b200.common.storage import Storage
    T, N, A, hw = 8, 32, 4, (14, 14)
    torch.manual_seed(4)
    pol = CategoricalPolicy(MLPModel(3 * 14 * 14, 4, 256, 64), False, A).to("cuda").flatten_()
    st = Storage((3, *hw), 64, T, N, "cuda")
    agent = PPO(None, pol, None, st, "cuda", 0, n_steps=T, n_envs=N, epoch=2, n_minibatch=2, mini_batch_size=128,
                learning_rate=5e-4, entropy_coef=0.01, matmul=matmul)
    if matmul == "tf32x3":
        assert isinstance(agent.engine, MLPEngineTC) and agent.engine.raw_pixels
    g = torch.Generator().manual_seed(1)
    frames = torch.randint(0, 256, (T + 1, N, *hw, 3), generator=g, dtype=torch.uint8)
    st.frames.copy_(frames.cuda())
    st.act_i32[:, :N] = torch.randint(0, A, (T, N), generator=g).cuda().int()
    st.logp[:, :N] = (-torch.rand(T, N, generator=g) * 2 - 0.5).cuda()
    st.value[:, :N] = (torch.randn(T + 1, N, generator=g) * 0.3).cuda()
    st.rew[:, :N] = torch.randn(T, N, generator=g).cuda()
    st.done_u8[:, :N] = (torch.rand(T, N, generator=g) < 0.1).cuda().to(torch.uint8)
    st.compute_estimates(0.999, 0.95, True, True)
    ref = oppo.OraclePolicy(oppo.OracleMLP(3 * 14 * 14, 4, 256, 64), A)
    ref.load_state_dict({k: v.detach().cpu() for k, v in pol.state_dict().items()})
    data = dict(obs=(frames[:-1].permute(0, 1, 4, 2, 3).float() / 255.0).reshape(T * N, -1),
                act=st.act_batch.cpu().reshape(-1), old_logp=st.log_prob_act_batch.cpu().reshape(-1),
                old_value=st.value_batch[:-1].cpu().reshape(-1), ret=st.return_batch.cpu().reshape(-1),
                adv=st.adv_batch.cpu().reshape(-1))
    opt = oppo.make_adam(ref, 5e-4)
    torch.manual_seed(99)
    logs = oppo.optimize(ref, opt, data, T, N, epoch=2, n_minibatch=2, mini_batch_size=128, grad_clip_norm=0.5,
                         eps_clip=0.2, value_coef=0.5, entropy_coef=0.01)
    torch.manual_seed(99)
    summary = agent.optimize()
    np.testing.assert_allclose(summary["Loss/total"], np.mean([l["total"] for l in logs]), rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(summary["Loss/entropy"], np.mean([l["entropy"] for l in logs]), rtol=2e-4)
    for (k, p), (_, q) in zip(pol.state_dict().items(), ref.state_dict().items()):
        np.testing.assert_allclose(p.cpu().numpy(), q.numpy(), rtol=5e-3, atol=2e-5, err_msg=k)
    # the predict() API hands the same policy float frames / 255: both first-layer weight copies must agree
    x = (frames[0].permute(0, 3, 1, 2).float() / 255.0).reshape(N, -1).cuda()
    head_f = agent._fwd(x, N).clone()
    head_r = agent._policy_head(st.obs_slot(0), st)
    torch.testing.assert_close(head_f[:, :A + 1], head_r[:, :A + 1], rtol=1e-4, atol=2e-6)
