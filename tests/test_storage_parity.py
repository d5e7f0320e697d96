"""GPU: GAE / normalisation / minibatch gather kernels (csrc/storage.cu) against the oracle and the fixtures."""
import os

import numpy as np
import pytest
import torch

from oracle import ppo as oppo

pytestmark = pytest.mark.gpu


def _storage(obs_shape, T, N):
    from tpp_b200.common.storage import Storage
    return Storage(obs_shape, 8, T, N, "cuda")


def test_gae_bit_exact_against_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "ppo.npz"))
    T, N = g["gae_rew"].shape
    st = _storage((9,), T, N)
    st.rew[:, :N] = torch.from_numpy(g["gae_rew"]).cuda()
    st.value[:, :N] = torch.from_numpy(g["gae_value"]).cuda()
    st.done_u8[:, :N] = torch.from_numpy(g["gae_done"]).cuda().to(torch.uint8)
    st.compute_estimates(0.99, 0.95, True, False)
    assert np.array_equal(st.adv_batch.cpu().numpy(), g["gae_adv_raw"])      # bit-exact (same fp32 op order)
    assert np.array_equal(st.return_batch.cpu().numpy(), g["gae_ret"])
    st.compute_estimates(0.99, 0.95, True, True)
    np.testing.assert_allclose(st.adv_batch.cpu().numpy(), g["gae_adv_norm"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("T,N", [(1, 2), (7, 5), (256, 256), (64, 4099), (256, 4096), (37, 8192), (16, 8200),
                                 (720, 100), (256, 65536)])
def test_gae_against_oracle(T, N):
    """Returns bit-exact on both kernels: the shared-memory staged scan (N <= 8192 and T*512 B <= 200 KB: the PPO
    configs) and the streaming thread-per-env scan (the C3 sweep sizes; long rollouts)."""
    gen = torch.Generator().manual_seed(T * 1000 + N)
    rew = torch.randn(T, N, generator=gen)
    value = torch.randn(T + 1, N, generator=gen)
    done = (torch.rand(T, N, generator=gen) < 0.05).float()
    st = _storage((3,), T, N)
    st.rew[:, :N], st.value[:, :N], st.done_u8[:, :N] = rew.cuda(), value.cuda(), done.cuda().to(torch.uint8)
    st.compute_estimates(0.999, 0.9, True, True)
    adv, ret = oppo.gae(rew, done, value, 0.999, 0.9)
    assert torch.equal(st.return_batch.cpu(), ret)
    if T * N > 1:
        np.testing.assert_allclose(st.adv_batch.cpu().numpy(), oppo.normalize_adv(adv).numpy(), rtol=1e-5, atol=2e-6)


@pytest.mark.parametrize("T,N", [(1, 2), (7, 5), (31, 33), (32, 64), (33, 40), (100, 300), (256, 256), (256, 4096),
                                 (257, 1000), (500, 2048), (256, 4736)])
@pytest.mark.parametrize("normalize", [True, False])
def test_gae_warp_scan_fused(T, N, normalize):
    """``gae_mode="warp_scan"``: warp-level segmented scan over n_steps + moments + normalisation in ONE launch
    (tpp_gae_scan).  The scan re-associates the recurrence's products, so it is compared within the stated tolerance
    (north_star: 1e-5 relative; here 1e-5 of the tensor's scale + 1e-5 relative) -- and against the exact kernels."""
    gen = torch.Generator().manual_seed(T * 1000 + N)
    rew = torch.randn(T, N, generator=gen)
    value = torch.randn(T + 1, N, generator=gen)
    done = (torch.rand(T, N, generator=gen) < 0.05).float()
    st = _storage((3,), T, N)
    st.gae_mode = "warp_scan"
    st.rew[:, :N], st.value[:, :N], st.done_u8[:, :N] = rew.cuda(), value.cuda(), done.cuda().to(torch.uint8)
    launches = st.n_launches
    st.compute_estimates(0.999, 0.9, True, normalize)
    assert st.n_launches - launches == 1, "the fused scan is one launch (normalisation included)"
    adv, ret = oppo.gae(rew, done, value, 0.999, 0.9)
    scale = float(adv.abs().max())
    np.testing.assert_allclose(st.return_batch.cpu().numpy(), ret.numpy(), rtol=1e-5, atol=1e-5 * scale)
    want = oppo.normalize_adv(adv) if (normalize and T * N > 1) else adv
    if normalize and T * N <= 1:
        return
    np.testing.assert_allclose(st.adv_batch.cpu().numpy(), want.numpy(), rtol=1e-5, atol=1e-5 * float(want.abs().max()))
    # moments of the raw advantages (what a sharded run all-reduces)
    m = st.moments.cpu().numpy()
    assert m[2] == T * N
    np.testing.assert_allclose(m[0], float(adv.double().sum()), rtol=1e-6, atol=1e-4 * scale)
    np.testing.assert_allclose(m[1], float((adv.double() ** 2).sum()), rtol=1e-5)


def test_gae_warp_scan_falls_back_outside_its_range():
    """N / 32 CTAs must be co-resident for the grid barrier: larger env counts run the exact kernels (bit-exact)."""
    T, N = 64, 1 << 16
    gen = torch.Generator().manual_seed(5)
    rew, value = torch.randn(T, N, generator=gen), torch.randn(T + 1, N, generator=gen)
    done = (torch.rand(T, N, generator=gen) < 0.05).float()
    st = _storage((3,), T, N)
    st.gae_mode = "warp_scan"
    st.rew[:, :N], st.value[:, :N], st.done_u8[:, :N] = rew.cuda(), value.cuda(), done.cuda().to(torch.uint8)
    st.compute_estimates(0.99, 0.95, True, True)
    assert torch.equal(st.return_batch.cpu(), oppo.gae(rew, done, value, 0.99, 0.95)[1])


def test_gae_warp_scan_inside_a_cuda_graph():
    """The cooperative launch is capturable: PPO.train's rollout graph ends with it."""
    T, N = 256, 4096
    gen = torch.Generator().manual_seed(9)
    rew, value = torch.randn(T, N, generator=gen), torch.randn(T + 1, N, generator=gen)
    done = (torch.rand(T, N, generator=gen) < 0.02).float()
    st = _storage((3,), T, N)
    st.gae_mode = "warp_scan"
    st.rew[:, :N], st.value[:, :N], st.done_u8[:, :N] = rew.cuda(), value.cuda(), done.cuda().to(torch.uint8)
    st.compute_estimates(0.99, 0.95, True, True)
    eager = st.adv_batch.clone()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            st.compute_estimates(0.99, 0.95, True, True)
    st.adv.zero_()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    np.testing.assert_allclose(st.adv_batch.cpu().numpy(), eager.cpu().numpy(), rtol=1e-6, atol=1e-6)


def test_use_gae_false_is_refused():
    st = _storage((3,), 4, 4)
    with pytest.raises(NotImplementedError):
        st.compute_estimates(use_gae=False)


def test_minibatch_indices_and_gather(golden_dir):
    """Same torch seed -> same index stream as the reference, and the gathered rows equal plain indexing."""
    g = np.load(os.path.join(golden_dir, "ppo.npz"))
    ref = g["mb_indices_seed1234_mb96"]
    T, N = 40, 24
    st = _storage((9,), T, N)
    gen = torch.Generator().manual_seed(0)
    obs = torch.randn(T + 1, N, 9, generator=gen)
    st.obs_batch[:] = obs.cuda()
    for name in ("logp", "value", "ret", "adv", "rew"):
        getattr(st, name)[:T, :N] = torch.randn(T, N, generator=gen).cuda()
    st.act_i32[:, :N] = torch.randint(0, 3, (T, N), generator=gen).cuda().int()
    st.done_u8[:, :N] = (torch.rand(T, N, generator=gen) < 0.1).cuda().to(torch.uint8)
    torch.manual_seed(1234)
    for e in range(ref.shape[0]):
        batches = list(st.fetch_train_generator(96))
        assert len(batches) == ref.shape[1]
        for i, (o, h, a, d, lp, v, r, adv) in enumerate(batches):
            idx = torch.from_numpy(ref[e, i])
            assert torch.equal(st.last_perm[i * 96:(i + 1) * 96], idx)
            assert torch.equal(o.cpu(), obs[:-1].reshape(T * N, 9)[idx])
            assert torch.equal(a.cpu(), st.act_batch.cpu().reshape(-1)[idx])
            assert torch.equal(d.cpu(), st.done_batch.cpu().reshape(-1)[idx])
            assert torch.equal(lp.cpu(), st.log_prob_act_batch.cpu().reshape(-1)[idx])
            assert torch.equal(v.cpu(), st.value_batch[:-1].cpu().reshape(-1)[idx])
            assert torch.equal(r.cpu(), st.return_batch.cpu().reshape(-1)[idx])
            assert torch.equal(adv.cpu(), st.adv_batch.cpu().reshape(-1)[idx])
            assert h.shape == (T * N, 8)


def test_image_gather_is_transpose_scale():
    T, N = 6, 8
    st = _storage((3, 14, 14), T, N)
    frames = torch.randint(0, 256, (T + 1, N, 14, 14, 3), dtype=torch.uint8)
    st.frames.copy_(frames.cuda())
    torch.manual_seed(5)
    for o, *_ in st.fetch_train_generator(16):
        idx = st.last_perm[:16]
        want = frames[:-1].reshape(T * N, 14, 14, 3)[idx].permute(0, 3, 1, 2).double() / 255.0
        np.testing.assert_allclose(o.cpu().double().numpy(), want.numpy(), rtol=0, atol=1e-7)
        break


def test_store_api_roundtrip():
    T, N = 3, 6
    st = _storage((5,), T, N)
    rng = np.random.default_rng(0)
    for t in range(T):
        st.store(rng.normal(size=(N, 5)), np.zeros((N, 8)), rng.integers(0, 3, N), rng.normal(size=N),
                 rng.random(N) < 0.3, [{} for _ in range(N)], rng.normal(size=N), rng.normal(size=N))
    st.store_last(rng.normal(size=(N, 5)), np.zeros((N, 8)), rng.normal(size=N))
    assert st.step == 0 and st.obs_batch.shape == (T + 1, N, 5) and st.act_batch.shape == (T, N)
    rew, done, _ = st.fetch_log_data()
    assert rew.shape == (T, N) and done.shape == (T, N)
