"""Logger parity: this package's Logger -- host ``feed`` and the device-side ``tpp_episode_scan`` ->
``feed_episodes`` path -- against the reference Logger (common/logger.py:55-174): every column of log-append.csv, the
episode deques and the episode count.  Fixture: tests/golden/logger.npz (minted from the live reference by
oracle/mint_golden.py); in the build container the unmodified reference Logger is also driven side by side."""
import contextlib
import csv
import io
import os
import warnings

import numpy as np
import pytest

from oracle.mint_golden import logger_batches
from oracle import ref_shim
from tpp_b200.common.logger import Logger, close_episodes

LOSS = ["Loss/pi", "Loss/v", "Loss/entropy", "Loss/x_entropy", "Loss/atn_entropy", "Loss/atn_entropy2",
        "Loss/sparsity", "Loss/feature_sparsity", "Loss/total"]


def _summary(i):
    return {k: 0.1 * (i + 1) * (j + 1) for j, k in enumerate(LOSS)}


def _rows(path):
    rows = list(csv.reader(open(path)))
    return rows[0], np.array([[float(x) if x != "" else np.nan for x in r] for r in rows[1:]])


def _check_rows(cols, got, want_cols, want):
    assert list(cols) == list(want_cols)
    skip = list(cols).index("wall_time")
    keep = [i for i in range(len(cols)) if i != skip]
    # float rewards: the reference sums an episode's float32 rewards with np.sum (pairwise, float32); here float64
    np.testing.assert_allclose(got[:, keep], want[:, keep], rtol=2e-6, atol=2e-6, equal_nan=True)


@pytest.mark.parametrize("tag,integer,max_steps", [("float", False, 7), ("int", True, 5)])
def test_host_feed_matches_reference_fixture(golden_dir, tmp_path, tag, integer, max_steps):
    g = np.load(os.path.join(golden_dir, "logger.npz"))
    lg = Logger(56, str(tmp_path))
    lg.max_steps = max_steps
    for i, (r, d, rv, dv) in enumerate(logger_batches(integer=integer)):
        lg.feed(r, d, np.nan, rv, dv, np.nan)
        lg.dump(_summary(i), 1e-3 / (i + 1))
    cols, rows = _rows(tmp_path / "log-append.csv")
    _check_rows(cols, rows, g[f"{tag}_columns"], g[f"{tag}_rows"])
    assert lg.num_episodes == int(g[f"{tag}_num_episodes"])
    np.testing.assert_allclose(np.array(lg.episode_reward_buffer), g[f"{tag}_reward_buffer"], rtol=2e-6, atol=2e-6)
    assert list(lg.episode_len_buffer) == list(g[f"{tag}_len_buffer"])
    assert list(lg.episode_timeout_buffer) == list(g[f"{tag}_timeout_buffer"])
    if integer:          # integer rewards (Box-World, cartpole): sums are exact in both
        assert np.array_equal(np.array(lg.episode_reward_buffer), g[f"{tag}_reward_buffer"])


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present (GPU box)")
def test_host_feed_matches_live_reference(tmp_path):
    ref_logger = ref_shim.load("common.logger")
    for seed, N, T, p in ((1, 7, 33, 0.2), (2, 300, 16, 0.01), (3, 64, 64, 0.5)):
        a, b = tmp_path / f"a{seed}", tmp_path / f"b{seed}"
        a.mkdir(); b.mkdir()
        ref, ours = ref_logger.Logger(N, str(a)), Logger(N, str(b))
        ref.max_steps = ours.max_steps = 3
        for i, (r, d, rv, dv) in enumerate(logger_batches(seed=seed, T=T, N=N, iters=4, p_done=p, integer=True)):
            ref.feed(r, d, np.nan, rv, dv, np.nan)
            ours.feed(r, d, np.nan, rv, dv, np.nan)
            with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
                warnings.simplefilter("ignore")
                ref.dump(_summary(i), 0.5)
            ours.dump(_summary(i), 0.5)
        ca, ra = _rows(a / "log-append.csv")
        cb, rb = _rows(b / "log-append.csv")
        _check_rows(cb, rb, ca, ra)
        assert list(ours.episode_len_buffer_v) == list(ref.episode_len_buffer_v)
        assert ours.num_episodes == ref.num_episodes


def test_close_episodes_is_the_reference_walk():
    """Brute-force restatement of common/logger.py:119-136 against the vectorised version."""
    rng = np.random.default_rng(0)
    N, T = 13, 29
    run_r, run_l = np.zeros(N), np.zeros(N, dtype=np.int64)
    open_eps = [[] for _ in range(N)]
    for _ in range(4):
        rew = rng.integers(-2, 3, (T, N)).astype(np.float64)
        done = rng.random((T, N)) < 0.15
        want_r, want_l = [], []
        for i in range(N):
            for j in range(T):
                open_eps[i].append(rew[j, i])
                if done[j, i]:
                    want_l.append(len(open_eps[i])); want_r.append(np.sum(open_eps[i])); open_eps[i] = []
        r, l = close_episodes(rew, done, run_r, run_l)
        assert np.array_equal(r, want_r) and np.array_equal(l, want_l)
        assert np.array_equal(run_l, [len(x) for x in open_eps])


@pytest.mark.gpu
@pytest.mark.parametrize("tag,integer,max_steps", [("float", False, 7), ("int", True, 5)])
def test_device_episode_scan_matches_reference_fixture(golden_dir, tmp_path, tag, integer, max_steps):
    """The device path end to end: rollout buffers in HBM -> tpp_episode_scan -> Logger.feed_episodes -> CSV rows equal
    to the reference Logger's."""
    import torch
    from tpp_b200.common.storage import Storage
    g = np.load(os.path.join(golden_dir, "logger.npz"))
    T, N = 48, 56
    st, sv = Storage((4,), 1, T, N, "cuda"), Storage((4,), 1, T, N, "cuda")
    lg = Logger(N, str(tmp_path))
    lg.max_steps = max_steps
    for i, (r, d, rv, dv) in enumerate(logger_batches(integer=integer)):
        for s, rr, dd in ((st, r, d), (sv, rv, dv)):
            s.rew[:, :N] = torch.from_numpy(rr).cuda()
            s.done_u8[:, :N] = torch.from_numpy(dd).cuda().to(torch.uint8)
        lg.feed_episodes(T, st.snapshot_episodes()(), np.nan, sv.snapshot_episodes()(), np.nan)
        lg.dump(_summary(i), 1e-3 / (i + 1))
    cols, rows = _rows(tmp_path / "log-append.csv")
    _check_rows(cols, rows, g[f"{tag}_columns"], g[f"{tag}_rows"])
    assert lg.num_episodes == int(g[f"{tag}_num_episodes"])
    assert list(lg.episode_len_buffer) == list(g[f"{tag}_len_buffer"])
    assert list(lg.episode_timeout_buffer_v) == [1 if l == max_steps else 0 for l in lg.episode_len_buffer_v]


@pytest.mark.gpu
@pytest.mark.parametrize("N,T,p", [(4096, 256, 0.02), (2, 5, 0.9), (70000, 16, 0.001), (33, 64, 0.0)])
def test_device_episode_scan_matches_host_walk(N, T, p):
    import torch
    from tpp_b200.common.storage import Storage
    rng = np.random.default_rng(N)
    st = Storage((1,), 1, T, N, "cuda")
    run_r, run_l = np.zeros(N), np.zeros(N, dtype=np.int64)
    for _ in range(3):
        rew = rng.integers(-3, 12, (T, N)).astype(np.float32)
        done = rng.random((T, N)) < p
        st.rew[:, :N] = torch.from_numpy(rew).cuda()
        st.done_u8[:, :N] = torch.from_numpy(done).cuda().to(torch.uint8)
        rec = st.snapshot_episodes()()
        r, l = close_episodes(rew, done, run_r, run_l)
        assert int(rec[0]) == len(r)
        kept = int(rec[1])
        assert kept == min(40, len(r))
        pairs = rec[2:2 + 2 * kept].reshape(kept, 2)
        assert np.array_equal(pairs[:, 0], r[len(r) - kept:]) and np.array_equal(pairs[:, 1], l[len(l) - kept:])
        assert np.array_equal(st._ep_state[0].cpu().numpy(), run_r)
        assert np.array_equal(st._ep_state[1].cpu().numpy(), run_l)
