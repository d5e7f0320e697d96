"""CPU: the algorithm of `tpp_gae_scan` (csrc/storage.cu, `gae_mode="warp_scan"`) restated in numpy with the kernel's
lane structure -- lane l composes its L = ceil(T / 32) time steps into one affine map, the 32 maps are combined by a
Kogge-Stone suffix scan, every lane replays its steps from its incoming A -- against the reference's sequential
recurrence (oracle.ppo.gae, common/storage.py:56-79).  This is where the tolerance stated for the scan form comes from:
the products are re-associated, the result agrees to a few 1e-7 of the advantage scale, never bit for bit."""
import numpy as np
import pytest
import torch

from oracle import ppo as oppo

f32 = np.float32


def scan_one_env(a, b):
    """A[t] = a[t] * A[t+1] + b[t], A[T] = 0, evaluated like one warp of the kernel does (all arithmetic in fp32)."""
    T = len(a)
    L = (T + 31) // 32
    P, Q = np.ones(32, f32), np.zeros(32, f32)
    for lane in range(32):
        p, q = f32(1), f32(0)
        for t in range(min(T, lane * L + L) - 1, lane * L - 1, -1):
            q = f32(a[t] * q + b[t])
            p = f32(a[t] * p)
        P[lane], Q[lane] = p, q
    d = 1
    while d < 32:                                   # inclusive suffix scan: S_l = f_l o f_{l+1} o ... o f_31
        P2, Q2 = np.ones(32, f32), np.zeros(32, f32)
        P2[:32 - d], Q2[:32 - d] = P[d:], Q[d:]
        for lane in range(32 - d):
            Q[lane] = f32(P[lane] * Q2[lane] + Q[lane])
            P[lane] = f32(P[lane] * P2[lane])
        d *= 2
    A = np.zeros(T, f32)
    for lane in range(32):
        x = Q[lane + 1] if lane < 31 else f32(0)    # A entering the lane's chunk = S_{l+1}(0)
        for t in range(min(T, lane * L + L) - 1, lane * L - 1, -1):
            x = f32(a[t] * x + b[t])
            A[t] = x
    return A


@pytest.mark.parametrize("T", [1, 5, 31, 32, 33, 100, 256, 257, 500])
def test_lane_structured_scan_equals_the_sequential_recurrence(T):
    gen = torch.Generator().manual_seed(T)
    N = 6
    rew, value = torch.randn(T, N, generator=gen), torch.randn(T + 1, N, generator=gen)
    done = (torch.rand(T, N, generator=gen) < 0.05).float()
    gamma, lmbda = 0.999, 0.9
    adv, ret = oppo.gae(rew, done, value, gamma, lmbda)
    gl = f32(np.float64(gamma) * np.float64(lmbda))
    nd = (1 - done.numpy()).astype(f32)
    delta = ((rew.numpy() + (f32(gamma) * value[1:].numpy()) * nd) - value[:-1].numpy()).astype(f32)
    scale = float(adv.abs().max())
    for e in range(N):
        A = scan_one_env((gl * nd[:, e]).astype(f32), delta[:, e])
        assert np.abs(A - adv[:, e].numpy()).max() <= 1e-6 * scale        # measured: <= 2e-7
        np.testing.assert_allclose(A + value[:-1, e].numpy(), ret[:, e].numpy(), rtol=1e-5, atol=1e-5 * scale)


def test_done_cuts_the_chain():
    """done = 1 makes a = 0: nothing crosses an episode boundary, whatever lane the boundary falls into."""
    T = 64
    a = np.full(T, 0.9, f32)
    b = np.ones(T, f32)
    a[40] = 0.0
    A = scan_one_env(a, b)
    tail = scan_one_env(a[41:], b[41:])
    assert A[40] == 1.0 and np.allclose(A[41:], tail, rtol=1e-6)
    head_only = scan_one_env(np.concatenate((a[:40], [f32(0)])), np.concatenate((b[:40], [f32(1)])))
    assert np.allclose(A[:41], head_only, rtol=1e-6)
