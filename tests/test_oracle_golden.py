"""CPU: the oracle (oracle/*.py) against the fixtures minted from the live reference (tests/golden/*.npz)."""
import os

import numpy as np
import pytest
import torch

from oracle import boxworld as obw
from oracle import ppo as oppo
from oracle.prevec import OraclePreVec

FAMILIES = ["cartpole", "cartpole_swing", "mountain_car", "acrobot"]


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def _kwargs(g):
    kw = {k: float(v) for k, v in zip(g["kwargs_keys"], g["kwargs_vals"])}
    kw["max_steps"] = int(kw["max_steps"])
    return kw


@pytest.mark.parametrize("family", FAMILIES)
def test_prevec_teacher_forced_matches_reference_bitwise(golden_dir, family):
    g = _load(golden_dir, f"prevec_{family}.npz")
    S, N = g["action"].shape
    env = OraclePreVec(family, N, seed=0, **_kwargs(g))
    for t in range(S):
        env.state = g["state_before"][t].copy()
        env.n_steps = g["n_steps_before"][t].copy()
        obs, rew, done = env.step(g["action"][t], reset_rows=g["reset_rows"][t])
        assert np.array_equal(done, g["done"][t])
        assert np.array_equal(rew, g["reward"][t])
        assert np.array_equal(obs, g["obs_after"][t])
        if family != "acrobot":
            assert np.array_equal(env.state, g["state_after"][t])
        else:   # reference wrap() couples envs: raw angles agree modulo 2*pi
            d = (env.state[:, :2] - g["state_after"][t][:, :2]) / (2 * np.pi)
            assert np.allclose(d, np.round(d), atol=1e-12)
            assert np.array_equal(env.state[:, 2:], g["state_after"][t][:, 2:])


@pytest.mark.parametrize("family", FAMILIES)
def test_prevec_free_running_same_pcg64_stream(golden_dir, family):
    """Seeded like the reference, the oracle reproduces its trajectory, resets included (same PCG64 draws)."""
    g = _load(golden_dir, f"prevec_{family}.npz")
    S, N = g["action"].shape
    env = OraclePreVec(family, N, seed=11, **_kwargs(g))
    env.reset()
    for t in range(S):
        obs, rew, done = env.step(g["action"][t])
        assert np.array_equal(done, g["done"][t]) and np.array_equal(obs, g["obs_after"][t]), (family, t)


def test_prevec_rejects_single_env():
    with pytest.raises(Exception):
        OraclePreVec("cartpole", 1)


@pytest.mark.parametrize("cfg", ["easy", "full", "mid"])
def test_boxworld_generator_matches_reference(golden_dir, cfg):
    g = _load(golden_dir, "boxworld.npz")
    spec = {"easy": (6, 2, 1, 1), "full": (12, 5, 3, 3), "mid": (12, 4, 2, 2)}[cfg]
    for i, seed in enumerate(g["gen_seeds"]):
        w, p, d = obw.world_gen(*spec, int(seed))
        assert np.array_equal(w, g[f"gen_{cfg}_world"][i])
        assert np.array_equal(p, g[f"gen_{cfg}_pos"][i])
        assert np.array_equal(d.astype(np.int8), g[f"gen_{cfg}_dic"][i])


@pytest.mark.parametrize("name,spec,n_levels", [("easy", (6, 2, 1, 1), 0), ("full", (12, 5, 3, 3), 0),
                                                ("easy_bank", (6, 2, 1, 1), 7), ("mid_bank", (12, 4, 2, 2), 500)])
def test_boxworld_trajectory_bit_exact(golden_dir, name, spec, n_levels):
    g = _load(golden_dir, "boxworld.npz")
    acts = g[f"traj_{name}_action"]
    env = obw.BoxWorldOracle(acts.shape[1], *spec, max_steps=25, start_seed=6033, n_levels=n_levels)
    assert np.array_equal(env.world, g[f"traj_{name}_world0"])
    for t in range(acts.shape[0]):
        w, r, d = env.step(acts[t])
        assert np.array_equal(w, g[f"traj_{name}_world"][t]), t
        assert np.array_equal(r, g[f"traj_{name}_reward"][t])
        assert np.array_equal(d, g[f"traj_{name}_done"][t])
        assert env.seed_counter == g[f"traj_{name}_seed_counter"][t]


SCENARIOS = {"keys_are_locked": [2, 2, 1, 1], "locks_dont_open_when_no_key": [1], "north_boundary": [0, 0, 0],
             "west_boundary": [3, 3, 3], "south_boundary": [3, 1, 1, 1, 1], "east_boundary": [0, 2, 2, 2, 2],
             "gem_inaccessible": [2, 1], "free_key_accessible": [2, 2, 2],
             "distractor_ends_game": [2, 2, 2, 1, 1, 1, 3, 3, 0], "goal_reachable": [2, 2, 2, 3, 3, 3, 1]}
IMPOSSIBLE = ["keys_are_locked", "locks_dont_open_when_no_key", "north_boundary", "west_boundary", "south_boundary",
              "east_boundary", "gem_inaccessible"]


@pytest.mark.parametrize("name", list(SCENARIOS))
def test_boxworld_reference_scenarios(golden_dir, name):
    """The reference's own ten known-answer tests (boxworld/box_world_env_vec_test.py:21-59) on the oracle."""
    g = _load(golden_dir, "boxworld.npz")
    env = obw.BoxWorldOracle(160, 6, 2, 1, 1, start_seed=0)
    env.replace_world_i(0, 0)
    before = None
    for a in SCENARIOS[name]:
        before = env.world[0].copy()
        w, r, d = env.step(np.full(160, a))
    assert np.array_equal(before, g[f"scn_{name}_before_last"])
    assert int(r[0]) == int(g[f"scn_{name}_reward"]) and bool(d[0]) == bool(g[f"scn_{name}_done"])
    if name in IMPOSSIBLE:
        assert np.array_equal(before, env.world[0])
    if name == "free_key_accessible":
        assert r[0] == 1
    if name == "distractor_ends_game":
        assert r[0] == -1 and d[0]
    if name == "goal_reachable":
        assert r[0] == 11 and d[0] and env.solved[0] and bool(g[f"scn_{name}_solved"])


def test_gae_and_normalisation_bit_exact(golden_dir):
    g = _load(golden_dir, "ppo.npz")
    adv, ret = oppo.gae(torch.from_numpy(g["gae_rew"]), torch.from_numpy(g["gae_done"]),
                        torch.from_numpy(g["gae_value"]), 0.99, 0.95)
    assert np.array_equal(adv.numpy(), g["gae_adv_raw"]) and np.array_equal(ret.numpy(), g["gae_ret"])
    assert np.array_equal(oppo.normalize_adv(adv).numpy(), g["gae_adv_norm"])


def test_minibatch_indices_bit_exact(golden_dir):
    g = _load(golden_dir, "ppo.npz")
    ref = g["mb_indices_seed1234_mb96"]
    torch.manual_seed(1234)
    for e in range(ref.shape[0]):
        got = np.array(oppo.epoch_indices(40 * 24, 96))
        assert np.array_equal(got, ref[e])


@pytest.mark.parametrize("tag,x_coef", [("plain", 0.0), ("xent", 0.05)])
def test_optimize_matches_reference(golden_dir, tag, x_coef):
    g = _load(golden_dir, "ppo.npz")
    T, N, A = 16, 16, 3
    pol = oppo.OraclePolicy(oppo.OracleMLP(9, 4, 32, 16), A)
    with torch.no_grad():
        for n, p in pol.named_parameters():
            p.copy_(torch.from_numpy(g[f"opt_{tag}_init/{n}"]))
    data = dict(obs=torch.from_numpy(g[f"opt_{tag}_obs_batch"][:-1]).reshape(T * N, 9),
                act=torch.from_numpy(g[f"opt_{tag}_act_batch"]).reshape(-1),
                old_logp=torch.from_numpy(g[f"opt_{tag}_log_prob_act_batch"]).reshape(-1),
                old_value=torch.from_numpy(g[f"opt_{tag}_value_batch"][:-1]).reshape(-1),
                ret=torch.from_numpy(g[f"opt_{tag}_return_batch"]).reshape(-1),
                adv=torch.from_numpy(g[f"opt_{tag}_adv_batch"]).reshape(-1))
    opt = oppo.make_adam(pol, 5e-3)
    torch.manual_seed(4321)
    logs = oppo.optimize(pol, opt, data, T, N, epoch=2, n_minibatch=4, mini_batch_size=64, grad_clip_norm=0.5,
                         eps_clip=0.2, value_coef=0.5, entropy_coef=0.02, x_entropy_coef=x_coef)
    for n, p in pol.named_parameters():
        np.testing.assert_allclose(p.detach().numpy(), g[f"opt_{tag}_final/{n}"], rtol=1e-6, atol=1e-8)
    summ = dict(zip(g[f"opt_{tag}_summary_keys"], g[f"opt_{tag}_summary_vals"]))
    np.testing.assert_allclose(np.mean([-l["pi_loss"] for l in logs]), summ["Loss/pi"], rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(np.mean([-l["value_loss"] for l in logs]), summ["Loss/v"], rtol=1e-6)
    np.testing.assert_allclose(np.mean([l["entropy"] for l in logs]), summ["Loss/entropy"], rtol=1e-6)
    np.testing.assert_allclose(np.mean([l["x_entropy"] for l in logs]), summ["Loss/x_entropy"], rtol=1e-5, atol=1e-9)
    np.testing.assert_allclose(np.mean([l["total"] for l in logs]), summ["Loss/total"], rtol=1e-6, atol=1e-9)


def test_recurrent_policy_chain_and_optimize_match_reference(golden_dir):
    """Row N4: the oracle's recurrent policy (GRU cell at prediction time) and its optimize() on env-permuting
    minibatches against fixtures minted from the live reference (oracle/mint_golden.py::mint_recurrent)."""
    g = _load(golden_dir, "recurrent.npz")
    T, N, A, D = 16, 16, 3, 64
    pol = oppo.OraclePolicy(oppo.OracleMLP(9, 4, 64, D), A, recurrent=True)
    assert [n for n, _ in pol.named_parameters()] == [str(n) for n in g["param_names"]]
    with torch.no_grad():
        for n, p in pol.named_parameters():
            p.copy_(torch.from_numpy(g[f"init/{n}"]))
        h = torch.zeros(N, D)
        for t in range(g["chain_obs"].shape[0]):
            dist, v, h = pol.predict(torch.from_numpy(g["chain_obs"][t]), h, 1 - torch.from_numpy(g["chain_done_prev"][t]))
            assert np.array_equal(dist.logits.numpy(), g["chain_logits"][t])
            assert np.array_equal(v.numpy(), g["chain_value"][t]) and np.array_equal(h.numpy(), g["chain_hidden"][t + 1])
    data = dict(obs=torch.from_numpy(g["opt_obs_batch"][:-1]).reshape(T * N, 9),
                act=torch.from_numpy(g["opt_act_batch"]).reshape(-1),
                old_logp=torch.from_numpy(g["opt_log_prob_act_batch"]).reshape(-1),
                old_value=torch.from_numpy(g["opt_value_batch"][:-1]).reshape(-1),
                ret=torch.from_numpy(g["opt_return_batch"]).reshape(-1), adv=torch.from_numpy(g["opt_adv_batch"]).reshape(-1))
    opt = oppo.make_adam(pol, 5e-3)
    torch.manual_seed(4321)
    logs = oppo.optimize(pol, opt, data, T, N, epoch=2, n_minibatch=4, mini_batch_size=64, grad_clip_norm=0.5,
                         eps_clip=0.2, value_coef=0.5, entropy_coef=0.02, x_entropy_coef=0.0, recurrent=True)
    for n, p in pol.named_parameters():
        np.testing.assert_allclose(p.detach().numpy(), g[f"final/{n}"], rtol=1e-6, atol=1e-8, err_msg=n)
        if n.startswith("gru."):                 # optimize() never calls the GRU: no gradient, no Adam state, no change
            assert np.array_equal(p.detach().numpy(), g[f"init/{n}"])
    summ = dict(zip(g["summary_keys"], g["summary_vals"]))
    np.testing.assert_allclose(np.mean([l["total"] for l in logs]), summ["Loss/total"], rtol=1e-6, atol=1e-9)
    assert len(g["adam_state_params"]) == len(g["param_names"]) - 4
