"""GPU: Box-World kernels (csrc/boxworld.cu) — integer state, bit-exact against the oracle and the fixtures."""
import os

import numpy as np
import pytest
import torch

from oracle import boxworld as obw

pytestmark = pytest.mark.gpu

SPECS = {"easy": (6, 2, 1, 1), "full": (12, 5, 3, 3), "mid": (12, 4, 2, 2)}


def _env(N, spec, **kw):
    from tpp_b200.boxworld.box_world_env_vec import BoxWorldVec
    return BoxWorldVec(N, *spec, **kw)


@pytest.mark.parametrize("cfg", list(SPECS))
def test_device_generator_matches_reference_levels(golden_dir, cfg):
    g = np.load(os.path.join(golden_dir, "boxworld.npz"))
    seeds = g["gen_seeds"]
    env = _env(len(seeds), SPECS[cfg])
    env._gen_into(None, seeds)
    assert np.array_equal(env.world.cpu().numpy(), g[f"gen_{cfg}_world"])
    assert np.array_equal(env.world_dic.cpu().numpy(), g[f"gen_{cfg}_dic"])
    assert np.array_equal(env.player_position.cpu().numpy(), g[f"gen_{cfg}_pos"])


@pytest.mark.parametrize("name,cfg,n_levels", [("easy", "easy", 0), ("full", "full", 0), ("easy_bank", "easy", 7),
                                               ("mid_bank", "mid", 500)])
def test_trajectory_bit_exact_against_golden(golden_dir, name, cfg, n_levels):
    g = np.load(os.path.join(golden_dir, "boxworld.npz"))
    acts = g[f"traj_{name}_action"]
    env = _env(acts.shape[1], SPECS[cfg], max_steps=25, start_seed=6033, n_levels=n_levels)
    assert np.array_equal(env.world.cpu().numpy(), g[f"traj_{name}_world0"])
    for t in range(acts.shape[0]):
        w, r, d, _ = env.step(acts[t])
        assert np.array_equal(w.cpu().numpy(), g[f"traj_{name}_world"][t]), t
        assert np.array_equal(r.cpu().numpy(), g[f"traj_{name}_reward"][t])
        assert np.array_equal(d.cpu().numpy(), g[f"traj_{name}_done"][t])
        assert env.np_random_seed == g[f"traj_{name}_seed_counter"][t]


SCENARIOS = {"keys_are_locked": [2, 2, 1, 1], "locks_dont_open_when_no_key": [1], "north_boundary": [0, 0, 0],
             "west_boundary": [3, 3, 3], "south_boundary": [3, 1, 1, 1, 1], "east_boundary": [0, 2, 2, 2, 2],
             "gem_inaccessible": [2, 1], "free_key_accessible": [2, 2, 2],
             "distractor_ends_game": [2, 2, 2, 1, 1, 1, 3, 3, 0], "goal_reachable": [2, 2, 2, 3, 3, 3, 1]}
IMPOSSIBLE = ["keys_are_locked", "locks_dont_open_when_no_key", "north_boundary", "west_boundary", "south_boundary",
              "east_boundary", "gem_inaccessible"]


@pytest.mark.parametrize("name", list(SCENARIOS))
def test_reference_scenarios(golden_dir, name):
    """The reference's ten known-answer tests (boxworld/box_world_env_vec_test.py:8-84) replayed on the kernel."""
    g = np.load(os.path.join(golden_dir, "boxworld.npz"))
    env = _env(160, (6, 2, 1, 1), start_seed=0)
    env.replace_world_i(0, 0)
    before = None
    for a in SCENARIOS[name]:
        before = env.world[0].cpu().numpy().copy()
        w, r, d, info = env.step(np.full(160, a))
    assert np.array_equal(before, g[f"scn_{name}_before_last"])
    assert int(r[0]) == int(g[f"scn_{name}_reward"]) and bool(d[0]) == bool(g[f"scn_{name}_done"])
    if name in IMPOSSIBLE:
        assert np.array_equal(before, env.world[0].cpu().numpy())
    if name == "free_key_accessible":
        assert int(r[0]) == 1
    if name == "distractor_ends_game":
        assert int(r[0]) == -1 and bool(d[0])
    if name == "goal_reachable":
        assert int(r[0]) == 11 and bool(d[0]) and info[0]["episode"]["solved"]


@pytest.mark.parametrize("N,spec,n_levels", [(4096, (12, 5, 3, 3), 500), (1000, (6, 2, 1, 1), 0), (3, (7, 3, 1, 2), 0),
                                             (16384, (6, 2, 1, 1), 50)])   # > 1024 CTAs: prefix-scan reset path
def test_long_trajectory_against_oracle(N, spec, n_levels):
    """Full-size config (N=4096, n=12, 500-level bank), odd grid (byte-copy path) and ragged N, with a policy
    that walks towards keys often enough to open locks; every array compared for equality at every step."""
    ms = 5 if N > 8192 else 30
    env = _env(N, spec, max_steps=ms, start_seed=7, n_levels=n_levels)
    orc = obw.BoxWorldOracle(N, *spec, max_steps=ms, start_seed=7, n_levels=n_levels)
    rng = np.random.default_rng(1)
    steps = (12 if N > 8192 else 40) if N > 2000 else 120
    for t in range(steps):
        a = rng.integers(0, 4, N)
        w, r, d, _ = env.step(a)
        ow, orr, od = orc.step(a)
        assert np.array_equal(w.cpu().numpy(), ow), t
        assert np.array_equal(r.cpu().numpy(), orr) and np.array_equal(d.cpu().numpy(), od)
        assert np.array_equal(env.player_position.cpu().numpy(), orc.player_position)
        assert np.array_equal(env.owned_key.cpu().numpy(), orc.owned_key)
        assert np.array_equal(env.num_env_steps.cpu().numpy(), orc.num_env_steps)
        assert np.array_equal(env.episode_reward.cpu().numpy(), orc.episode_reward)
        assert env.np_random_seed == orc.seed_counter
    assert np.array_equal(env.world_dic.cpu().numpy(), orc.world_dic.astype(np.int8))


def test_wrapped_env_matches_oracle_wrappers():
    """VecNormalize(ob=False) + TransposeFrame + ScaledFloatFrame semantics (procgen_wrappers.py:314-419)."""
    from tpp_b200.boxworld.box_world_env_vec import create_bw_env
    hp = dict(n_envs=64, grid_size=6, goal_length=2, num_distractor=1, distractor_length=1, max_steps=20)

    class A:
        seed, num_levels = 3, 11
    venv = create_bw_env(A, hp)
    assert venv.observation_space.shape == (3, 8, 8)
    orc = obw.BoxWorldOracle(64, 6, 2, 1, 1, max_steps=20, start_seed=3, n_levels=11)
    vn = obw.VecNormalizeOracle(64)
    rng = np.random.default_rng(0)
    obs = venv.reset()
    np.testing.assert_allclose(obs.cpu().numpy(), obw.frame_to_obs(orc.world), atol=1e-7)
    for t in range(60):
        a = rng.integers(0, 4, 64)
        obs, rew, done, info = venv.step(a)
        ow, orr, od = orc.step(a)
        want = vn.step(orr.astype(np.float64), od)
        np.testing.assert_allclose(obs.cpu().numpy(), obw.frame_to_obs(ow), atol=1e-7)
        np.testing.assert_allclose(rew.cpu().numpy(), want, rtol=1e-6, atol=1e-7)
        assert np.array_equal(done.cpu().numpy(), od)
    assert info[0]["env_reward"] == int(orr[0])
