"""GPU: recurrent policies.  CategoricalPolicy(recurrent=True): the GRU cell acts in predict() /
rollouts (csrc/gru.cu + two tensor-core GEMMs), optimize() runs embedder + heads on env-permuting minibatches without
it -- exactly what the reference does (agents/ppo.py:72-81, 116-121; common/storage.py:93-110).  Checked against torch's
nn.GRU, the oracle, and fixtures minted from the live reference (tests/golden/recurrent.npz)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = dict(rtol=2e-5, atol=2e-5)      # forward quantities of the 3xTF32 engine


def _policy(in_dim, A, mid, latent, seed=0, recurrent=True):
    from tpp_b200.common.model import MLPModel
    from tpp_b200.common.policy import CategoricalPolicy
    torch.manual_seed(seed)
    return CategoricalPolicy(MLPModel(in_dim, 4, mid, latent), recurrent, A)


@pytest.mark.parametrize("N,D", [(7, 64), (300, 64), (1000, 256), (4096, 256)])
def test_gru_cell_matches_torch_gru(N, D):
    from tpp_b200 import _lib
    from tpp_b200.common.engine import GRUCellTC
    pol = _policy(9, 3, 32, D, seed=N).to("cuda").flatten_()
    cell = GRUCellTC(pol)
    g = torch.Generator(device="cuda").manual_seed(D + N)
    x = torch.randn(N, D, device="cuda", generator=g)
    h = torch.randn(N, D, device="cuda", generator=g)
    done = (torch.rand(N, device="cuda", generator=g) < 0.3).to(torch.uint8)
    ld = cell.ld
    hi, lo = torch.zeros(N, ld, device="cuda"), torch.zeros(N, ld, device="cuda")
    _lib.call("tpp_split_tf32", _lib.ptr(x), D, N, D, _lib.ptr(hi), _lib.ptr(lo), ld, None, None, 0, _lib.stream_ptr())
    h_out = torch.zeros(N, D, device="cuda")
    (o_hi, o_lo), ld_o = cell.step((hi, lo), ld, h, done, h_out, N)
    ref = torch.nn.GRU(D, D)          # torch CPU fp32 (cuDNN's GRU would run its GEMMs in TF32)
    ref.load_state_dict({k.split(".")[-1]: v.cpu() for k, v in pol.state_dict().items() if k.startswith("gru.gru.")})
    with torch.no_grad():
        want = ref(x.cpu()[None], (h.cpu() * (1.0 - done.float().cpu())[:, None])[None])[1][0].cuda()
    torch.testing.assert_close(h_out, want, **TOL)
    torch.testing.assert_close((o_hi + o_lo)[:, :D], h_out, rtol=0, atol=1e-6)
    assert (o_hi[:, D:] == 0).all() and (o_lo[:, D:] == 0).all()
    # in place (the bootstrap step) and without a mask
    h2 = h.clone()
    cell.step((hi, lo), ld, h2, None, h2, N)
    with torch.no_grad():
        want2 = ref(x.cpu()[None], h.cpu()[None])[1][0].cuda()
    torch.testing.assert_close(h2, want2, **TOL)


def _agent(pol, obs_shape, T, N, **kw):
    from tpp_b200.agents.ppo import PPO
    from tpp_b200.common.storage import Storage
    st = Storage(obs_shape, pol.embedder.output_dim, T, N, "cuda")
    return PPO(kw.pop("env", None), pol, kw.pop("logger", None), st, "cuda", 0, n_steps=T, n_envs=N, **kw), st


def test_predict_chain_matches_reference_fixture(golden_dir):
    """PPO.predict(obs, hidden_state, done) over a chain of steps: logits are not returned, so the check is value,
    next hidden state and log-prob of the drawn action against the reference's recorded logits."""
    g = np.load(os.path.join(golden_dir, "recurrent.npz"))
    N, A, D = 16, 3, 64
    pol = _policy(9, A, 64, D)
    pol.load_state_dict({str(n): torch.from_numpy(g[f"init/{n}"]) for n in g["param_names"]})
    pol = pol.to("cuda").flatten_()
    agent, _ = _agent(pol, (9,), 4, N)
    h = np.zeros((N, D), dtype=np.float32)
    for t in range(g["chain_obs"].shape[0]):
        act, logp, value, h = agent.predict(g["chain_obs"][t], h, g["chain_done_prev"][t])
        np.testing.assert_allclose(h, g["chain_hidden"][t + 1], **TOL)
        np.testing.assert_allclose(value, g["chain_value"][t], **TOL)
        np.testing.assert_allclose(logp, g["chain_logits"][t][np.arange(N), act], **TOL)
        h = g["chain_hidden"][t + 1]          # teacher-forced: errors do not compound across the chain


def test_optimize_matches_reference_fixture_and_leaves_gru_untouched(golden_dir):
    g = np.load(os.path.join(golden_dir, "recurrent.npz"))
    T, N, A, D = 16, 16, 3, 64
    pol = _policy(9, A, 64, D)
    pol.load_state_dict({str(n): torch.from_numpy(g[f"init/{n}"]) for n in g["param_names"]})
    pol = pol.to("cuda").flatten_()
    agent, st = _agent(pol, (9,), T, N, epoch=2, n_minibatch=4, mini_batch_size=64, gamma=0.99, lmbda=0.95,
                       learning_rate=5e-3, grad_clip_norm=0.5, eps_clip=0.2, value_coef=0.5, entropy_coef=0.02)
    st.obs_batch[:] = torch.from_numpy(g["opt_obs_batch"]).cuda()
    st.act_i32[:, :N] = torch.from_numpy(g["opt_act_batch"]).cuda().int()
    st.logp[:, :N] = torch.from_numpy(g["opt_log_prob_act_batch"]).cuda()
    st.value[:, :N] = torch.from_numpy(g["opt_value_batch"]).cuda()
    st.rew[:, :N] = torch.from_numpy(g["opt_rew_batch"]).cuda()
    st.done_u8[:, :N] = torch.from_numpy(g["opt_done_batch"]).cuda().to(torch.uint8)
    st.compute_estimates(0.99, 0.95, True, True)
    gru0 = {n: p.detach().clone() for n, p in pol.named_parameters() if n.startswith("gru.")}
    torch.manual_seed(4321)
    summary = agent.optimize()
    want = dict(zip([str(k) for k in g["summary_keys"]], g["summary_vals"]))
    for k in ("Loss/pi", "Loss/v", "Loss/entropy", "Loss/total"):
        np.testing.assert_allclose(summary[k], want[k], rtol=2e-4, atol=2e-6, err_msg=k)
    for n, p in pol.named_parameters():
        np.testing.assert_allclose(p.detach().cpu().numpy(), g[f"final/{n}"], rtol=2e-4, atol=5e-6, err_msg=n)
    for n, p0 in gru0.items():
        assert torch.equal(dict(pol.named_parameters())[n].detach(), p0), n      # zero gradient => Adam moves nothing
    assert agent.optimizer.step_count == int(g["adam_step"])


def test_fetch_train_generator_recurrent_yields_whole_trajectories():
    from tpp_b200.common.storage import Storage
    T, N, D = 8, 12, 5
    st = Storage((3,), D, T, N, "cuda")
    code = torch.arange((T + 1) * N, dtype=torch.float32, device="cuda").view(T + 1, N)
    st.obs_batch[:] = code[:, :, None].expand(T + 1, N, 3)
    st.hidden_states_batch[:] = code[:, :, None] + torch.arange(D, device="cuda") / 8.0
    st.adv[:, :N] = -code[:T]
    torch.manual_seed(9)
    batches = list(st.fetch_train_generator(mini_batch_size=24, recurrent=True))
    torch.manual_seed(9)
    perm = torch.randperm(N).view(4, 3)
    assert len(batches) == 4
    for (obs, hid, act, done, logp, value, ret, adv), envs in zip(batches, perm):
        want = (torch.arange(T)[:, None] * N + envs[None, :]).reshape(-1).float().cuda()
        assert torch.equal(obs[:, 0], want) and torch.equal(adv, -want)
        assert torch.equal(hid, st.hidden_states_batch[0, envs.cuda()])


@pytest.mark.parametrize("use_graph", [False, True])
def test_device_rollout_hidden_chain_matches_oracle(use_graph):
    """Two rollouts on a device env: every slot's hidden state, value and log-prob re-derived with the oracle's
    predict() from the stored observations / actions / done flags -- including the reference's quirks: slot T holds the
    state AFTER the bootstrap predict, and the next rollout starts from it with the last step's done mask."""
    from oracle import ppo as oppo
    from tpp_b200.discrete_env.cartpole_pre_vec import CartPoleVecEnv
    T, N, A, D = 12, 64, 2, 64
    env = CartPoleVecEnv(n_envs=N, seed=3, max_steps=9, device="cuda")
    pol = _policy(env.observation_space.shape[0], A, 64, D, seed=5).to("cuda")
    orc = oppo.OraclePolicy(oppo.OracleMLP(env.observation_space.shape[0], 4, 64, D), A, recurrent=True)
    orc.load_state_dict({k: v.cpu() for k, v in pol.state_dict().items()})
    pol = pol.flatten_()
    agent, st = _agent(pol, env.observation_space.shape, T, N, env=env, use_cuda_graph=use_graph)
    env.reset_rollout(st)
    h_prev_T, done_carry = torch.zeros(N, D), torch.zeros(N)
    for it in range(3):
        agent.collect_rollout(env, st)
        torch.cuda.synchronize()
        obs = st.obs_batch.cpu()
        hid = st.hidden_states_batch.cpu()
        done = st.done_u8[:, :N].float().cpu()
        act = st.act_i32[:, :N].long().cpu()
        torch.testing.assert_close(hid[0], h_prev_T, rtol=0, atol=0)         # carried over bit for bit
        with torch.no_grad():
            for t in range(T + 1):
                mask = 1 - (done_carry if t == 0 else done[t - 1])
                dist, v, h_next = orc.predict(obs[t], hid[t] if t < T else h_T_in, mask)
                if t < T:
                    if t + 1 < T:
                        torch.testing.assert_close(hid[t + 1], h_next, **TOL)
                    else:
                        h_T_in = h_next                      # slot T was then advanced in place by the bootstrap call
                    torch.testing.assert_close(st.value[t, :N].cpu(), v, **TOL)
                    torch.testing.assert_close(st.logp[t, :N].cpu(), dist.logits[torch.arange(N), act[t]], **TOL)
                else:
                    torch.testing.assert_close(st.value[T, :N].cpu(), v, **TOL)
                    torch.testing.assert_close(hid[T], h_next, **TOL)
        assert done.sum() > 0                                # masks were exercised
        h_prev_T, done_carry = hid[T].clone(), done[T - 1].clone()
        agent._carry_over(st)


def test_train_recurrent_device_and_host_envs(tmp_path):
    """PPO.train end to end with a recurrent policy: device env (graphs on) and a host-stepped numpy env."""
    from tpp_b200.discrete_env.cartpole_pre_vec import CartPoleVecEnv
    T, N, A, D = 16, 32, 2, 64
    env = CartPoleVecEnv(n_envs=N, seed=1, max_steps=20, device="cuda")
    pol = _policy(env.observation_space.shape[0], A, 64, D, seed=2).to("cuda").flatten_()
    agent, st = _agent(pol, env.observation_space.shape, T, N, env=env, epoch=2, n_minibatch=2, mini_batch_size=256)
    gru0 = pol.flat[pol.layout["gru.gru.weight_ih_l0"][0]:].clone()
    w0 = pol.flat[:1000].clone()
    agent.train(T * N * 3)
    assert torch.isfinite(pol.flat).all() and not torch.equal(pol.flat[:1000], w0)
    assert torch.equal(pol.flat[pol.layout["gru.gru.weight_ih_l0"][0]:], gru0)
    assert st.hidden_states_batch.abs().sum() > 0

    class HostEnv:
        def __init__(self, n):
            self.n, self.rng, self.t = n, np.random.default_rng(0), 0

        def reset(self):
            return self.rng.normal(size=(self.n, 4)).astype(np.float32)

        def step(self, act):
            self.t += 1
            done = self.rng.random(self.n) < 0.1
            return self.reset(), self.rng.normal(size=self.n).astype(np.float32), done, [{} for _ in range(self.n)]

        def close(self):
            pass

    pol2 = _policy(4, A, 64, D, seed=3).to("cuda").flatten_()
    agent2, st2 = _agent(pol2, (4,), T, N, env=HostEnv(N), epoch=1, n_minibatch=2, mini_batch_size=256)
    agent2.train(T * N * 3)
    assert torch.isfinite(pol2.flat).all()
    assert st2.hidden_states_batch.abs().sum() > 0 and st2.done_u8.sum() > 0


def test_impala_recurrent_predict_matches_oracle():
    """The same cell behind the IMPALA embedder (hidden size 256): one predict step against the oracle."""
    from oracle import ppo as oppo
    from tpp_b200.agents.ppo import PPO
    from tpp_b200.common.model import ImpalaModel
    from tpp_b200.common.policy import CategoricalPolicy
    from tpp_b200.common.storage import Storage
    N, A = 24, 15
    torch.manual_seed(8)
    pol = CategoricalPolicy(ImpalaModel(3), True, A)
    orc = oppo.OraclePolicy(oppo.OracleImpala(3), A, recurrent=True)
    orc.load_state_dict(pol.state_dict())
    pol = pol.to("cuda").flatten_()
    st = Storage((3, 64, 64), 256, 4, N, "cuda")
    agent = PPO(None, pol, None, st, "cuda", 0, n_steps=4, n_envs=N)
    g = torch.Generator().manual_seed(1)
    obs = torch.randint(0, 256, (N, 3, 64, 64), generator=g).float() / 255.0
    h = torch.randn(N, 256, generator=g) * 0.5
    done = (torch.rand(N, generator=g) < 0.4).float()
    act, logp, value, h_next = agent.predict(obs.numpy(), h.numpy(), done.numpy())
    with torch.no_grad():
        dist, v, h_want = orc.predict(obs, h, 1 - done)
    # IMPALA forward tolerance (heads 3e-5 of the output scale)
    np.testing.assert_allclose(h_next, h_want.numpy(), rtol=1e-4, atol=5e-5)
    np.testing.assert_allclose(value, v.numpy(), rtol=1e-4, atol=5e-5)
    np.testing.assert_allclose(logp, dist.logits.numpy()[np.arange(N), act], rtol=1e-4, atol=5e-5)
