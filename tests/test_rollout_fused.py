"""GPU: the fused rollout-step policy kernel (csrc/rollout_fused.cu, SURVEY 8f N1: 4 dense layers + heads + action draw
in one cluster launch, K-split over 4 CTAs with a DSMEM reduce-scatter) against (a) a float64 torch evaluation of the
same MLPModel + heads (reference common/model.py:954-980, common/policy.py:74-87) and (b) the per-layer GEMM path +
tpp_sample_actions it replaces.  Stated tolerance: logits / value within 2e-5 of the float64 result relative to the
output scale (3xTF32, fp32-grade); the draw is the same Philox inverse-CDF, so given the kernel's own logits the action
and log-prob must equal tpp_sample_actions' exactly."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _policy(in_dim, A, raw):
    from tpp_b200.common.engine import MLPEngineTC
    from tpp_b200.common.model import MLPModel
    from tpp_b200.common.policy import CategoricalPolicy
    torch.manual_seed(11)
    pol = CategoricalPolicy(MLPModel(in_dim, 4, 256, 64), False, A).to("cuda").flatten_()
    with torch.no_grad():                       # non-trivial biases / head scale so that every term is exercised
        pol.flat.add_(0.02 * torch.randn_like(pol.flat))
        pol.fc_policy.weight.mul_(30.0)
    eng = MLPEngineTC(pol, A, raw_pixels=raw)
    return pol, eng


def _ref_head(pol, x64):
    sd = {k: v.detach().double().cpu() for k, v in pol.state_dict().items()}
    h = x64
    lin = sorted({k.rsplit(".", 1)[0] for k in sd if k.startswith("embedder.")}, key=lambda s: [int(t) if t.isdigit() else t
                                                                                              for t in s.split(".")])
    for i, name in enumerate(lin):
        h = h @ sd[name + ".weight"].t() + sd[name + ".bias"]
        if i < len(lin) - 1:
            h = h.clamp_min(0)
    logits = h @ sd["fc_policy.weight"].t() + sd["fc_policy.bias"]
    value = h @ sd["fc_value.weight"].t() + sd["fc_value.bias"]
    return torch.cat((logits, value), 1)


def _check(head, ref, tol=2e-5):
    err = (head.double().cpu() - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= tol * scale, f"max err {err:.3e} vs scale {scale:.3e}"


@pytest.mark.parametrize("N", [4096, 300, 128 * 40 + 5])
def test_fused_rollout_policy_on_pixel_rows(N):
    """Box-World shape: raw pixel rows [N][608] (exact TF32 operand, 1/255 in the first layer's weight copy)."""
    from tpp_b200 import _lib
    A, in_dim = 4, 588
    pol, eng = _policy(in_dim, A, raw=True)
    assert eng.fused_rollout_ok(True)
    g = torch.Generator().manual_seed(N)
    px = torch.zeros(N, eng.ld_in)
    px[:, :in_dim] = torch.randint(0, 256, (N, in_dim), generator=g).float()
    px = px.cuda()
    act, logp, value = (torch.zeros(N, dtype=torch.int32, device="cuda"), torch.zeros(N, device="cuda"),
                        torch.zeros(N, device="cuda"))
    head = torch.zeros(N, eng.ld_head, device="cuda")
    tick = torch.full((1,), 7, dtype=torch.int64, device="cuda")
    eng.rollout_fused(px, N, eng.ld_in, True, act, logp, value, 123, tick, 5, env_offset=3, head_out=head)
    torch.cuda.synchronize()
    _check(head[:, :A + 1], _ref_head(pol, px[:, :in_dim].double().cpu() / 255.0))
    # the per-layer GEMM path it replaces
    old = eng.forward(px, N, raw=True, need_backward=False)
    _check(head[:, :A + 1], old[:, :A + 1].double().cpu())
    # same draw as tpp_sample_actions on the same logits
    a2, l2, v2 = torch.zeros_like(act), torch.zeros_like(logp), torch.zeros_like(value)
    _lib.call("tpp_sample_actions", _lib.ptr(head), eng.ld_head, N, A, _lib.ptr(a2), _lib.ptr(l2), _lib.ptr(v2), 123,
              _lib.ptr(tick), 5, 0, 3, _lib.stream_ptr())
    assert torch.equal(act, a2) and torch.equal(logp, l2) and torch.equal(value, v2)
    assert len(torch.unique(act)) == A and (logp <= 0).all()


@pytest.mark.parametrize("N", [4096, 77])
def test_fused_rollout_policy_on_uint8_frames(N):
    """The uint8 NHWC frames themselves as the first-layer operand (weight columns in frame byte order): same logits as
    the float NCHW / 255 contract of TransposeFrame + ScaledFloatFrame (common/env/procgen_wrappers.py:391-419)."""
    from tpp_b200.common.engine import MLPEngineTC
    from tpp_b200.common.model import MLPModel
    from tpp_b200.common.policy import CategoricalPolicy
    A, shape = 4, (3, 14, 14)
    torch.manual_seed(12)
    pol = CategoricalPolicy(MLPModel(588, 4, 256, 64), False, A).to("cuda").flatten_()
    with torch.no_grad():
        pol.flat.add_(0.02 * torch.randn_like(pol.flat))
        pol.fc_policy.weight.mul_(30.0)
    eng = MLPEngineTC(pol, A, raw_pixels=True, obs_shape=shape)
    assert eng.w0_bytes is not None
    g = torch.Generator().manual_seed(N)
    frames = torch.randint(0, 256, (N, 14, 14, 3), generator=g, dtype=torch.uint8).cuda()
    act, logp, value = (torch.zeros(N, dtype=torch.int32, device="cuda"), torch.zeros(N, device="cuda"),
                        torch.zeros(N, device="cuda"))
    head = torch.zeros(N, eng.ld_head, device="cuda")
    tick = torch.zeros(1, dtype=torch.int64, device="cuda")
    eng.rollout_fused(frames, N, 588, "u8", act, logp, value, 1, tick, 0, head_out=head)
    torch.cuda.synchronize()
    x = frames.permute(0, 3, 1, 2).reshape(N, -1).double().cpu() / 255.0        # NCHW flatten / 255
    _check(head[:, :A + 1], _ref_head(pol, x))


@pytest.mark.parametrize("N,n_obs,A", [(256, 9, 2), (4096, 14, 3), (100, 5, 3)])
def test_fused_rollout_policy_on_feature_major_slots(N, n_obs, A):
    """Vector envs (cartpole / acrobot / mountain car): the layer-1 operand is the feature-major rollout slot."""
    pol, eng = _policy(n_obs, A, raw=False)
    assert eng.fused_rollout_ok(False)
    ld = (N + 31) // 32 * 32
    x = torch.zeros(n_obs, ld, device="cuda")
    x[:, :N] = torch.randn(n_obs, N, device="cuda") * 2.0
    act, logp, value = (torch.zeros(N, dtype=torch.int32, device="cuda"), torch.zeros(N, device="cuda"),
                        torch.zeros(N, device="cuda"))
    head = torch.zeros(N, eng.ld_head, device="cuda")
    tick = torch.zeros(1, dtype=torch.int64, device="cuda")
    eng.rollout_fused(x, N, ld, False, act, logp, value, 0, tick, 0, head_out=head)
    torch.cuda.synchronize()
    ref = _ref_head(pol, x[:, :N].t().double().cpu())
    _check(head[:, :A + 1], ref)
    lp = torch.log_softmax(ref[:, :A], 1).gather(1, act.long().cpu()[:, None])[:, 0]
    np.testing.assert_allclose(logp.cpu().numpy(), lp.numpy(), rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(value.cpu().numpy(), ref[:, A].numpy(), rtol=1e-4, atol=2e-5 * ref.abs().max().item())


def test_fused_and_unfused_rollouts_train_alike():
    """PPO.train on Box-World with the fused step kernel vs the per-layer path: same seeds -> same env trajectories as
    long as the sampled actions agree (logits differ by ~1e-6, so a handful of near-tie draws may flip)."""
    from tpp_b200.agents.ppo import PPO
    from tpp_b200.boxworld.box_world_env_vec import create_bw_env
    from tpp_b200.common.model import MLPModel
    from tpp_b200.common.policy import CategoricalPolicy
    from tpp_b200.common.storage import Storage
    outs = []
    for fused in (True, False):
        hp = dict(n_envs=512, grid_size=12, goal_length=5, num_distractor=3, distractor_length=3, max_steps=40)
        env = create_bw_env(None, hp)
        torch.manual_seed(6033)
        pol = CategoricalPolicy(MLPModel(588, 4, 256, 64), False, 4).to("cuda").flatten_()
        st = Storage((3, 14, 14), 64, 16, 512, "cuda")
        agent = PPO(env, pol, None, st, "cuda", 0, n_steps=16, n_envs=512, epoch=1, n_minibatch=2, mini_batch_size=4096,
                    fused_rollout=fused)
        env.reset_rollout(st)
        agent.collect_rollout(env, st)         # eager
        outs.append((st.act_i32.clone(), st.logp.clone(), st.value.clone(), st.frames.clone()))
    same = (outs[0][0][0] == outs[1][0][0]).float().mean().item()
    assert same > 0.999                                       # step 0: identical inputs
    torch.testing.assert_close(outs[0][2][0], outs[1][2][0], rtol=1e-4, atol=1e-5)
    m = outs[0][0][0] == outs[1][0][0]
    torch.testing.assert_close(outs[0][1][0][m], outs[1][1][0][m], rtol=1e-4, atol=1e-5)
