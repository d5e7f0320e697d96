"""GPU: policy engine, fused PPO loss fwd+bwd, clip+Adam and the whole optimize() against the oracle / fixtures.

Tolerance (stated): fp32 everywhere; loss terms, gradients and updated parameters must agree with the torch-CPU
oracle to |a-b| <= 1e-6 + 1e-5*|b| unless noted (summation order differs: CUDA reductions vs torch CPU)."""
import os

import numpy as np
import pytest
import torch

from oracle import ppo as oppo

pytestmark = pytest.mark.gpu


def _policy(in_dim, A, depth=4, mid=32, latent=16, seed=0):
    from tpp_b200.common.model import MLPModel
    from tpp_b200.common.policy import CategoricalPolicy
    torch.manual_seed(seed)
    pol = CategoricalPolicy(MLPModel(in_dim, depth, mid, latent), False, A)
    return pol.to("cuda").flatten_()


def _close(a, b, rtol=1e-5, atol=1e-6, what=""):
    np.testing.assert_allclose(np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64), rtol=rtol, atol=atol,
                               err_msg=what)


@pytest.mark.parametrize("M,in_dim,A,mid,latent", [(64, 9, 2, 32, 16), (4096, 9, 2, 256, 64), (1000, 588, 4, 256, 64),
                                                   (130, 5, 3, 48, 24), (256, 14, 15, 256, 64)])
@pytest.mark.parametrize("kind", ["fp32", "tf32x3", "tf32x3-split-on-chip"])
def test_engine_forward_backward_vs_torch(M, in_dim, A, mid, latent, kind):
    """Hand-written forward/backward on the flat buffer == torch autograd on the same (aliased) parameters, for
    the CUDA-core fp32 engine and the tcgen05 3xTF32 engine (both must hold the fp32 tolerance); the latter also with
    the hidden activations kept as plain fp32 arrays whose lo halves are formed on chip (TPP_TC_A_SPLIT / B_SPLIT)."""
    from tpp_b200.common.engine import MLPEngine, MLPEngineTC
    torch.backends.cuda.matmul.allow_tf32 = False
    pol = _policy(in_dim, A, mid=mid, latent=latent, seed=M)
    eng = MLPEngine(pol, A) if kind == "fp32" else MLPEngineTC(pol, A, precision=3, split_on_chip=kind.endswith("chip"))
    if kind.endswith("chip") and not eng.split_on_chip:
        pytest.skip("a layer narrower than 33 columns: the engine keeps (hi, lo) pairs")
    x = torch.randn(M, (in_dim + 3) // 4 * 4, device="cuda")[:, :in_dim]
    head = eng.forward(x, M)
    dist, v, _ = pol(x, None, None)
    logits = pol.fc_policy(pol.embedder(x))
    # two independent fp32-grade evaluations of a 5-layer network (cuBLAS fp32 vs ours): rounding of the K<=588
    # contractions compounds to ~1e-5 absolute on O(1) outputs -> stated atol 2e-5 (+ 1e-5 relative)
    _close(head[:, :A].cpu(), logits.detach().cpu(), atol=2e-5, what="logits")
    _close(head[:, A].cpu(), v.detach().cpu(), atol=2e-5, what="value")
    dhead = torch.zeros_like(head)
    dhead[:, :A + 1] = torch.randn(M, A + 1, device="cuda") / M
    pol.flat_grad.zero_()
    eng.backward(dhead, M)
    mine = pol.flat_grad.clone()
    pol.flat_grad.zero_()
    out = torch.cat((logits, pol.fc_value(pol.embedder(x))), 1)
    out.backward(dhead[:, :A + 1])
    ref = pol.flat_grad.clone()
    scale = ref.abs().max().item()
    _close(mine.cpu(), ref.cpu(), rtol=1e-4, atol=2e-6 * max(scale, 1.0), what="flat gradient")


@pytest.mark.parametrize("kind", ["fp32", "tf32x3"])
def test_engine_feature_major_input_equals_row_major(kind):
    from tpp_b200.common.engine import MLPEngine, MLPEngineTC
    pol = _policy(9, 2, mid=64, latent=32)
    eng = MLPEngine(pol, 2) if kind == "fp32" else MLPEngineTC(pol, 2)
    N, ld = 300, 320
    x = torch.randn(N, 9, device="cuda")
    fm = torch.zeros(9, ld, device="cuda")
    fm[:, :N] = x.t()
    xp = torch.zeros(N, 12, device="cuda")
    xp[:, :9] = x
    a = eng.forward(xp[:, :9], N).clone()
    b = eng.forward(fm, N, feature_major_ld=ld).clone()
    if kind == "fp32":
        assert torch.equal(a, b)
    else:   # first layer runs on CUDA cores for feature-major input and on tensor cores for row-major input
        _close(a.cpu(), b.cpu(), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("A,x_coef", [(2, 0.0), (3, 0.05), (15, 0.0), (4, 0.3)])
def test_loss_kernel_vs_oracle_autograd(A, x_coef):
    import ctypes as C
    from tpp_b200 import _lib
    B = 1000
    g = torch.Generator().manual_seed(A)
    logits = torch.randn(B, A, generator=g) * 1.5
    value = torch.randn(B, generator=g)
    act = torch.randint(0, A, (B,), generator=g)
    old_logp = torch.log_softmax(logits + 0.4 * torch.randn(B, A, generator=g), 1).gather(1, act[:, None])[:, 0]
    old_value = value + 0.3 * torch.randn(B, generator=g)
    ret = value + torch.randn(B, generator=g)
    adv = torch.randn(B, generator=g)
    adv[::17] = 0.0
    old_value[::13] = value[::13]                   # exercise the tie branches of clamp / max
    lt, vt = logits.clone().requires_grad_(), value.clone().requires_grad_()
    dist = torch.distributions.Categorical(logits=torch.log_softmax(lt, 1))
    loss, terms = oppo.ppo_loss(dist, vt, act.float(), old_logp, old_value, ret, adv, 0.2, 0.5, 0.02, 0.7, x_coef)
    loss.backward()
    ld = (A + 1 + 3) // 4 * 4
    head = torch.zeros(B, ld)
    head[:, :A], head[:, A] = logits, value
    head = head.cuda()
    dhead = torch.full((B, ld), 7.0, device="cuda")
    stats = torch.zeros(20, dtype=torch.float64, device="cuda")
    pbar = torch.zeros(16, device="cuda")
    cfg = _lib.LossCfg(0.2, 0.5, 0.02, 0.7, x_coef, A, B)
    s = _lib.stream_ptr()
    _lib.call("tpp_ppo_pbar", _lib.ptr(head), ld, B, A, _lib.ptr(pbar), s)
    dev = lambda t, dt=torch.float32: t.to("cuda", dt)
    bufs = (dev(act, torch.int32), dev(old_logp), dev(old_value), dev(ret), dev(adv))
    _lib.call("tpp_ppo_loss_fwd_bwd", C.byref(cfg), _lib.ptr(head), ld, *[_lib.ptr(b) for b in bufs],
              _lib.ptr(pbar) if x_coef else None, _lib.ptr(dhead), _lib.ptr(stats), s)
    S = stats.cpu().numpy()
    _close(-S[0] / B, terms["pi_loss"].item(), what="pi_loss")
    _close(0.5 * S[1] / B, terms["value_loss"].item(), what="value_loss")
    _close(S[2] / B, terms["entropy"].item(), what="entropy")
    p = S[4:4 + A] / B
    _close(-(p * np.log(p)).sum() - S[2] / B, terms["x_entropy"].item(), atol=2e-6, what="x_entropy")
    _close(dhead[:, :A].cpu(), lt.grad, rtol=1e-4, atol=1e-8, what="dlogits")
    _close(dhead[:, A].cpu(), vt.grad, rtol=1e-5, atol=1e-9, what="dvalue")
    assert (dhead[:, A + 1:] == 0).all()


def test_adam_clip_vs_torch():
    """clip_grad_norm_(0.5) + Adam(eps=1e-5) with a changing lr, five steps, both sides of the clip threshold."""
    from tpp_b200.agents.ppo import FlatAdam
    pol = _policy(9, 2, mid=64, latent=32)
    ref = [p.detach().clone().cpu().requires_grad_() for p in pol.parameters()]
    opt_ref = torch.optim.Adam(ref, lr=3e-3, eps=1e-5)
    opt = FlatAdam(pol, 3e-3, eps=1e-5, max_grad_norm=0.5)
    g = torch.Generator().manual_seed(0)
    for step in range(5):
        scale = [10.0, 0.01, 1.0, 3.0, 1e-3][step]
        lr = 3e-3 * (1 - step / 10)
        opt.param_groups[0]["lr"] = lr
        for gp in opt_ref.param_groups:
            gp["lr"] = lr
        for p_ref, p in zip(ref, pol.parameters()):
            gr = torch.randn(p_ref.shape, generator=g) * scale
            p_ref.grad = gr.clone()
            p.grad.copy_(gr.cuda())
        torch.nn.utils.clip_grad_norm_(ref, 0.5)
        opt_ref.step()
        opt.step()
        assert float(pol.flat_grad.abs().max()) == 0.0       # gradients zeroed by the fused kernel
    assert opt.step_count == 5
    for p_ref, p in zip(ref, pol.parameters()):
        _close(p.detach().cpu(), p_ref.detach(), rtol=1e-5, atol=1e-7, what="param after 5 Adam steps")
    sd, sd_ref = opt.state_dict(), opt_ref.state_dict()
    _close(sd["state"][0]["exp_avg"].cpu(), sd_ref["state"][0]["exp_avg"], rtol=1e-5, atol=1e-9)
    _close(sd["state"][0]["exp_avg_sq"].cpu(), sd_ref["state"][0]["exp_avg_sq"], rtol=1e-5, atol=1e-12)


@pytest.mark.parametrize("matmul", ["fp32", "tf32x3"])
@pytest.mark.parametrize("tag,x_coef", [("plain", 0.0), ("xent", 0.05)])
def test_optimize_against_reference_fixture(golden_dir, tag, x_coef, matmul):
    """PPO.optimize on the reference's recorded rollout + initial weights + torch seed: same minibatches, and
    final parameters / logged summary within fp32 tolerance of what the reference produced."""
    from tpp_b200.agents.ppo import PPO
    from tpp_b200.common.model import MLPModel
    from tpp_b200.common.policy import CategoricalPolicy
    from tpp_b200.common.storage import Storage
    g = np.load(os.path.join(golden_dir, "ppo.npz"))
    T, N, A = 16, 16, 3
    pol = CategoricalPolicy(MLPModel(9, 4, 32, 16), False, A)
    pol.load_state_dict({str(n): torch.from_numpy(g[f"opt_{tag}_init/{n}"]) for n in g[f"opt_{tag}_param_names"]})
    pol = pol.to("cuda").flatten_()
    st = Storage((9,), 16, T, N, "cuda")
    st.obs_batch[:] = torch.from_numpy(g[f"opt_{tag}_obs_batch"]).cuda()
    st.act_i32[:, :N] = torch.from_numpy(g[f"opt_{tag}_act_batch"]).cuda().int()
    st.logp[:, :N] = torch.from_numpy(g[f"opt_{tag}_log_prob_act_batch"]).cuda()
    st.value[:, :N] = torch.from_numpy(g[f"opt_{tag}_value_batch"]).cuda()
    st.rew[:, :N] = torch.from_numpy(g[f"opt_{tag}_rew_batch"]).cuda()
    st.done_u8[:, :N] = torch.from_numpy(g[f"opt_{tag}_done_batch"]).cuda().to(torch.uint8)
    st.compute_estimates(0.99, 0.95, True, True)
    _close(st.return_batch.cpu(), g[f"opt_{tag}_return_batch"], what="returns")
    _close(st.adv_batch.cpu(), g[f"opt_{tag}_adv_batch"], atol=2e-6, what="normalised advantages")
    agent = PPO(None, pol, None, st, "cuda", 1, n_steps=T, n_envs=N, epoch=2, n_minibatch=4, mini_batch_size=64,
                gamma=0.99, lmbda=0.95, learning_rate=5e-3, grad_clip_norm=0.5, eps_clip=0.2, value_coef=0.5,
                entropy_coef=0.02, x_entropy_coef=x_coef, matmul=matmul)
    torch.manual_seed(4321)
    summary = agent.optimize()
    assert list(summary.keys()) == [str(k) for k in g[f"opt_{tag}_summary_keys"]]
    want = dict(zip(summary.keys(), g[f"opt_{tag}_summary_vals"]))
    for k in ("Loss/pi", "Loss/v", "Loss/entropy", "Loss/x_entropy", "Loss/total"):
        _close(summary[k], want[k], rtol=2e-4, atol=2e-6, what=k)
    for k in ("Loss/atn_entropy", "Loss/atn_entropy2", "Loss/sparsity", "Loss/feature_sparsity"):
        assert np.isnan(summary[k]) and np.isnan(want[k])
    for n, p in pol.named_parameters():
        _close(p.detach().cpu(), g[f"opt_{tag}_final/{n}"], rtol=2e-4, atol=5e-6, what=f"final {n}")
    assert agent.optimizer.step_count == int(g[f"opt_{tag}_adam_step"])


def test_predict_and_sampling_distribution():
    from tpp_b200.agents.ppo import PPO
    from tpp_b200.common.storage import Storage
    pol = _policy(9, 3, mid=64, latent=32, seed=3)
    with torch.no_grad():
        pol.fc_policy.bias.copy_(torch.tensor([0.5, -0.2, 0.1]))
    st = Storage((9,), 32, 4, 50000, "cuda")
    agent = PPO(None, pol, None, st, "cuda", 0, n_steps=4, n_envs=50000)
    obs = np.zeros((50000, 9), dtype=np.float64)
    act, logp, value, _ = agent.predict(obs, None, np.zeros(50000))
    assert act.dtype == np.int64 and act.shape == (50000,) and logp.shape == (50000,)
    dist, v, _ = pol(torch.zeros(1, 9, device="cuda"), None, None)
    p = dist.probs[0].detach().cpu().numpy()
    freq = np.bincount(act, minlength=3) / 50000
    assert np.abs(freq - p).max() < 0.01                      # Philox inverse-CDF draws follow the categorical
    _close(logp, np.log(p)[act], rtol=1e-5, atol=1e-6)
    _close(value, np.full(50000, v.item()), rtol=1e-5, atol=1e-6)
    act2, *_ = agent.predict(obs, None, np.zeros(50000))
    assert (act2 != act).any()                                # the tick advances the stream between calls


@pytest.mark.gpu
@pytest.mark.parametrize("matmul", ["fp32", "tf32x3"])
@pytest.mark.parametrize("image", [False, True])
def test_accumulation_groups_match_oracle_and_unfused(matmul, image):
    """Gradient accumulation (agents/ppo.py:111,173-177): the minibatches between two optimizer steps share one
    gather / forward / loss / backward pass (``fuse_accum``).  Per-minibatch loss terms and the final parameters
    must match the torch-CPU oracle, which runs them one by one, and the unfused engine path."""
    from oracle import ppo as oppo
    from tpp_b200.agents.ppo import PPO
    from tpp_b200.common.model import MLPModel
    from tpp_b200.common.policy import CategoricalPolicy
    from tpp_b200.common.storage import Storage
    T, N, A = 16, 128, 4
    obs_shape = (3, 6, 6) if image else (9,)
    in_dim = int(np.prod(obs_shape))
    kw = dict(n_steps=T, n_envs=N, epoch=2, n_minibatch=2, mini_batch_size=256, learning_rate=5e-4, entropy_coef=0.01,
              gamma=0.99, lmbda=0.95, matmul=matmul)   # batch_size 1024 -> 4 minibatches accumulate per step
    g = torch.Generator().manual_seed(7)
    frames = torch.randint(0, 256, (T + 1, N, *obs_shape[1:], 3), generator=g, dtype=torch.uint8) if image else None
    vec = torch.randn(T + 1, N, 9, generator=g)
    rec = dict(act=torch.randint(0, A, (T, N), generator=g).int(), logp=-torch.rand(T, N, generator=g) * 2 - 0.5,
               value=torch.randn(T + 1, N, generator=g) * 0.3, rew=torch.randn(T, N, generator=g),
               done=(torch.rand(T, N, generator=g) < 0.1).to(torch.uint8))
    results = {}
    for fuse in ("auto", 2, 1, "per-group graphs"):
        torch.manual_seed(4)
        pol = CategoricalPolicy(MLPModel(in_dim, 4, 64, 32), False, A).to("cuda").flatten_()
        init = {k: v.detach().cpu().clone() for k, v in pol.state_dict().items()}
        st = Storage(obs_shape, 32, T, N, "cuda")
        if fuse == "per-group graphs":         # the path multi-GPU runs take (all-reduce between group and step)
            agent = PPO(None, pol, None, st, "cuda", 0, fuse_accum="auto", epoch_graph=False, **kw)
        else:
            agent = PPO(None, pol, None, st, "cuda", 0, fuse_accum=fuse, **kw)
        if image:
            st.frames.copy_(frames.cuda())
        else:
            st.obs_batch[:] = vec.cuda()
        st.act_i32[:, :N], st.logp[:, :N], st.value[:, :N] = rec["act"].cuda(), rec["logp"].cuda(), rec["value"].cuda()
        st.rew[:, :N], st.done_u8[:, :N] = rec["rew"].cuda(), rec["done"].cuda()
        st.compute_estimates(0.99, 0.95, True, True)
        torch.manual_seed(99)
        summary = agent.optimize()
        assert agent._group_size(4, 8, 256, agent.engine) == {"auto": 4, 2: 2, 1: 1, "per-group graphs": 4}[fuse]
        assert agent.optimizer.step_count == 2 * 2
        results[fuse] = (summary, {k: v.detach().cpu() for k, v in pol.state_dict().items()},
                         {k: np.array(v) for k, v in agent.last_stats.items()})
    ref = oppo.OraclePolicy(oppo.OracleMLP(in_dim, 4, 64, 32), A)
    ref.load_state_dict(init)
    obs = (frames[:-1].permute(0, 1, 4, 2, 3).float() / 255.0) if image else vec[:-1]
    data = dict(obs=obs.reshape(T * N, -1), act=st.act_batch.cpu().reshape(-1),
                old_logp=st.log_prob_act_batch.cpu().reshape(-1), old_value=st.value_batch[:-1].cpu().reshape(-1),
                ret=st.return_batch.cpu().reshape(-1), adv=st.adv_batch.cpu().reshape(-1))
    opt = oppo.make_adam(ref, 5e-4)
    torch.manual_seed(99)
    logs = oppo.optimize(ref, opt, data, T, N, epoch=2, n_minibatch=2, mini_batch_size=256, grad_clip_norm=0.5,
                         eps_clip=0.2, value_coef=0.5, entropy_coef=0.01)
    assert len(logs) == 16                 # 8 minibatches per epoch, 4 per optimizer step
    for fuse, (summary, params, stats) in results.items():
        for key in ("pi_loss", "value_loss", "entropy", "total"):      # one entry per minibatch, in order
            np.testing.assert_allclose(stats[key], [l[key] for l in logs], rtol=2e-4, atol=2e-5,
                                       err_msg=f"fuse={fuse} {key}")
        for (k, p), (_, q) in zip(params.items(), ref.state_dict().items()):
            np.testing.assert_allclose(p.numpy(), q.numpy(), rtol=5e-3, atol=2e-5, err_msg=f"fuse={fuse} {k}")
    for k in results[1][1]:                                            # fused vs unfused: only summation order differs
        np.testing.assert_allclose(results["auto"][1][k].numpy(), results[1][1][k].numpy(), rtol=1e-4, atol=2e-6,
                                   err_msg=k)


@pytest.mark.gpu
@pytest.mark.parametrize("image,N", [(False, 256), (True, 1000)])
def test_rollout_tail_kernel_matches_gemm_head_and_sampler(image, N):
    """Rollout step, policy side: trunk GEMMs + tpp_mlp_tail_sample (last layer, heads and the draw in one CUDA-core
    launch) against the all-GEMM forward + tpp_sample_actions on the same slot: same draws, same log-probs / values."""
    from tpp_b200.agents.ppo import PPO
    from tpp_b200.common.engine import MLPEngineTC
    from tpp_b200.common.model import MLPModel
    from tpp_b200.common.policy import CategoricalPolicy
    from tpp_b200.common.storage import Storage
    T, A = 4, 4
    obs_shape = (3, 14, 14) if image else (9,)
    torch.manual_seed(11)
    pol = CategoricalPolicy(MLPModel(int(np.prod(obs_shape)), 4, 256, 64), False, A).to("cuda").flatten_()
    with torch.no_grad():
        pol.flat.mul_(3.0)                       # spread the logits so that the draws are not all near-uniform
    st = Storage(obs_shape, 64, T, N, "cuda")
    agent = PPO(None, pol, None, st, "cuda", 0, n_steps=T, n_envs=N, epoch=1, n_minibatch=1, mini_batch_size=N * T)
    assert isinstance(agent.engine, MLPEngineTC) and agent.engine.tail_ok()
    agent.engine.refresh_weights()
    g = torch.Generator().manual_seed(5)
    if image:
        st.frames.copy_(torch.randint(0, 256, st.frames.shape, generator=g, dtype=torch.uint8).cuda())
    else:
        st.obs_batch[:] = torch.randn(T + 1, N, 9, generator=g).cuda()
    outs = []
    for fused in (True, False):
        agent.fused_tail = fused
        st.act_i32.fill_(-1)
        agent._policy_sample(st, 2)
        outs.append((st.act_i32[2, :N].clone(), st.logp[2, :N].clone(), st.value[2, :N].clone()))
    (a1, l1, v1), (a0, l0, v0) = outs
    assert (a1 >= 0).all() and (a1 < A).all() and len(a1.unique()) == A
    assert (a1 == a0).float().mean() > 0.998          # a draw within ~1e-6 of a CDF step may fall either side
    same = a1 == a0
    torch.testing.assert_close(l1[same], l0[same], rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(v1, v0, rtol=1e-4, atol=1e-5 * float(v0.abs().max()) + 2e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("image", [False, True])
def test_adam_step_refreshes_operand_views(image):
    """tpp_adam_clip_step_views leaves every tensor-core operand copy of the parameters (plain, 1/255-scaled and
    frame-byte-order first layer, heads) exactly as tpp_split_tf32-based refresh_weights() would."""
    from tpp_b200.agents.ppo import FlatAdam
    from tpp_b200.common.engine import MLPEngineTC
    from tpp_b200.common.model import MLPModel
    from tpp_b200.common.policy import CategoricalPolicy
    torch.manual_seed(5)
    in_dim = 588 if image else 9
    pol = CategoricalPolicy(MLPModel(in_dim, 4, 256, 64), False, 4).to("cuda").flatten_()
    eng = MLPEngineTC(pol, 4, raw_pixels=image, obs_shape=(3, 14, 14) if image else None)
    opt = FlatAdam(pol, 1e-2)
    pairs = [w for w in eng.w] + [eng.wh] + ([eng.w0_raw, eng.w0_bytes] if image else [])
    for it in range(3):
        pol.flat_grad.copy_(torch.randn_like(pol.flat_grad))
        opt.step(eng.weight_views())
        got = [(w["hi"].clone(), w["lo"].clone()) for w in pairs]
        eng.refresh_weights()
        for (hi, lo), w in zip(got, pairs):
            assert torch.equal(hi, w["hi"]) and torch.equal(lo, w["lo"])
    assert opt.step_count == 3
