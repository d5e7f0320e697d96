"""Mint tests/golden/*.npz from the LIVE reference (run in the build container, where /root/reference exists):

    python -m oracle.mint_golden

The reference publishes no golden vectors for this path except Box-World's ten scripted scenarios
(boxworld/box_world_env_vec_test.py), so the fixtures are generated here by importing the unmodified reference
modules through oracle/ref_shim.py.  The GPU box has no /root/reference: tests there only read the .npz files.
Test infrastructure; never imported by the product package.
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

FAMILIES = {
    "cartpole": ("discrete_env.cartpole_pre_vec", "CartPoleVecEnv", 2, dict(max_steps=25)),
    "cartpole_swing": ("discrete_env.cartpole_swing_pre_vec", "CartPoleSwingVecEnv", 2, dict(max_steps=40)),
    "mountain_car": ("discrete_env.mountain_car_pre_vec", "MountainCarVecEnv", 3,
                     dict(max_steps=30, min_goal_position=-0.45, max_goal_position=0.6)),
    "acrobot": ("discrete_env.acrobot_pre_vec", "AcrobotVecEnv", 3, dict(max_steps=60)),
}


def mint_prevec(family, n_envs=48, steps=100, seed=11):
    mod, cls, n_act, kw = FAMILIES[family]
    env = getattr(ref_shim.load(mod), cls)(n_envs=n_envs, seed=seed, **kw)
    blocks = []
    orig = env.start_space.sample

    def recording_sample(n):
        b = orig(n)
        blocks.append(b.copy())
        return b

    env.start_space.sample = recording_sample
    env.reset()
    init_rows = blocks[-1].copy()
    rng = np.random.default_rng(seed)
    rec = {k: [] for k in ("state_before", "n_steps_before", "action", "reset_rows", "state_after", "obs_after",
                           "reward", "done")}
    for _ in range(steps):
        a = rng.integers(0, n_act, n_envs)
        rec["state_before"].append(env.state.copy())
        rec["n_steps_before"].append(env.n_steps.copy())
        nb = len(blocks)
        obs, rew, done, _ = env.step(a)
        rec["action"].append(a)
        rec["reset_rows"].append(blocks[-1].copy() if len(blocks) > nb else np.zeros_like(init_rows))
        rec["state_after"].append(env.state.copy())
        rec["obs_after"].append(np.asarray(obs).copy())
        rec["reward"].append(np.asarray(rew, dtype=np.float64).copy())
        rec["done"].append(np.asarray(done).copy())
    out = {k: np.stack(v) for k, v in rec.items()}
    out["init_rows"] = init_rows
    out["kwargs_keys"] = np.array(list(kw.keys()))
    out["kwargs_vals"] = np.array([float(v) for v in kw.values()])
    np.savez_compressed(os.path.join(OUT, f"prevec_{family}.npz"), **out)
    print(family, "dones:", int(out["done"].sum()))


BW_CONFIGS = {"easy": (6, 2, 1, 1), "full": (12, 5, 3, 3), "mid": (12, 4, 2, 2)}
SCENARIOS = {  # boxworld/box_world_env_vec_test.py:21-59
    "keys_are_locked": [2, 2, 1, 1], "locks_dont_open_when_no_key": [1], "north_boundary": [0, 0, 0],
    "west_boundary": [3, 3, 3], "south_boundary": [3, 1, 1, 1, 1], "east_boundary": [0, 2, 2, 2, 2],
    "gem_inaccessible": [2, 1], "free_key_accessible": [2, 2, 2],
    "distractor_ends_game": [2, 2, 2, 1, 1, 1, 3, 3, 0], "goal_reachable": [2, 2, 2, 3, 3, 3, 1],
}


def mint_boxworld():
    bw = ref_shim.load("boxworld.box_world_env_vec")
    gen = ref_shim.load("boxworld.boxworld_gen_vec")
    out = {}
    # generator: worlds / dic / pos for a spread of seeds
    seeds = np.array(list(range(24)) + [499, 500, 6033, 10 ** 6 + 5, 2 ** 31 + 9, 2 ** 40 + 3], dtype=np.int64)
    out["gen_seeds"] = seeds
    for name, cfg in BW_CONFIGS.items():
        ws, ps, ds = zip(*(gen.world_gen(*cfg, int(s)) for s in seeds))
        out[f"gen_{name}_world"] = np.concatenate(ws)
        out[f"gen_{name}_pos"] = np.concatenate(ps)
        out[f"gen_{name}_dic"] = np.concatenate(ds).astype(np.int8)
    # trajectories (random actions, short episodes so that resets and the seed counter are exercised)
    for name, cfg, n_levels in (("easy", BW_CONFIGS["easy"], 0), ("full", BW_CONFIGS["full"], 0),
                                ("easy_bank", BW_CONFIGS["easy"], 7), ("mid_bank", BW_CONFIGS["mid"], 500)):
        N, S = 32, 80
        env = bw.BoxWorldVec(N, *cfg, max_steps=25, start_seed=6033, n_levels=n_levels)
        rng = np.random.default_rng(5)
        acts, worlds, rews, dones, seeds_after = [], [], [], [], []
        w0 = env.world.copy()
        for _ in range(S):
            a = rng.integers(0, 4, N)
            w, r, d, _ = env.step(a)
            acts.append(a); worlds.append(w.copy()); rews.append(r.copy()); dones.append(d.copy())
            seeds_after.append(env.np_random_seed)
        out[f"traj_{name}_world0"] = w0
        out[f"traj_{name}_action"] = np.stack(acts)
        out[f"traj_{name}_world"] = np.stack(worlds)
        out[f"traj_{name}_reward"] = np.stack(rews)
        out[f"traj_{name}_done"] = np.stack(dones)
        out[f"traj_{name}_seed_counter"] = np.array(seeds_after)
        print("boxworld", name, "dones", int(np.stack(dones).sum()), "rewards!=0", int((np.stack(rews) != 0).sum()))
    # the reference's ten scripted scenarios on its own fixture (160 envs, 6/2/1/1, seed 0)
    env = bw.BoxWorldVec(160, 6, 2, 1, 1, start_seed=0)
    for name, script in SCENARIOS.items():
        env.replace_world_i(0, 0)
        w_prev = env.world[0].copy()
        for a in script:
            w_prev = env.world[0].copy()
            w, r, d, info = env.step(np.full(160, a))
        out[f"scn_{name}_before_last"] = w_prev
        out[f"scn_{name}_after_last"] = env.world[0].copy() if not d[0] else w_prev * 0
        out[f"scn_{name}_reward"] = np.array(r[0])
        out[f"scn_{name}_done"] = np.array(d[0])
        out[f"scn_{name}_solved"] = np.array(bool(info[0].get("episode", {}).get("solved", False)))
    np.savez_compressed(os.path.join(OUT, "boxworld.npz"), **out)


class _DummyLogger:
    episode_reward_buffer = [0.0]


def mint_ppo():
    storage_mod = ref_shim.load("common.storage")
    model_mod = ref_shim.load("common.model")
    policy_mod = ref_shim.load("common.policy")
    ppo_mod = ref_shim.load("agents.ppo")
    out = {}
    # --- GAE + normalisation + minibatch index stream ---
    T, N = 40, 24
    g = torch.Generator().manual_seed(3)
    st = storage_mod.Storage((9,), 4, T, N, "cpu")
    st.rew_batch = torch.randn(T, N, generator=g)
    st.value_batch = torch.randn(T + 1, N, generator=g)
    st.done_batch = (torch.rand(T, N, generator=g) < 0.08).float()
    st.obs_batch[:-1, :, 0] = torch.arange(T * N, dtype=torch.float32).view(T, N)
    out["gae_rew"], out["gae_value"], out["gae_done"] = st.rew_batch.numpy(), st.value_batch.numpy(), \
        st.done_batch.numpy()
    st.compute_estimates(0.99, 0.95, True, False)
    out["gae_adv_raw"], out["gae_ret"] = st.adv_batch.numpy().copy(), st.return_batch.numpy().copy()
    st.compute_estimates(0.99, 0.95, True, True)
    out["gae_adv_norm"] = st.adv_batch.numpy().copy()
    torch.manual_seed(1234)
    epochs = []
    for _ in range(3):
        epochs.append(np.stack([s[0][:, 0].numpy().astype(np.int64) for s in st.fetch_train_generator(96)]))
    out["mb_indices_seed1234_mb96"] = np.stack(epochs)

    # --- full optimize() on a fixed rollout: loss terms, gradients, clip, Adam ---
    for tag, x_coef in (("plain", 0.0), ("xent", 0.05)):
        T, N, A = 16, 16, 3
        torch.manual_seed(77)
        emb = model_mod.MLPModel(in_channels=9, depth=4, mid_weight=32, latent_size=16)
        pol = policy_mod.CategoricalPolicy(emb, False, A)
        names = [n for n, _ in pol.named_parameters()]
        for n, p in pol.named_parameters():
            out[f"opt_{tag}_init/{n}"] = p.detach().numpy().copy()
        st = storage_mod.Storage((9,), 16, T, N, "cpu")
        g = torch.Generator().manual_seed(5)
        st.obs_batch = torch.randn(T + 1, N, 9, generator=g)
        st.act_batch = torch.randint(0, A, (T, N), generator=g).float()
        st.log_prob_act_batch = -torch.rand(T, N, generator=g) * 1.5 - 0.4
        st.value_batch = torch.randn(T + 1, N, generator=g) * 0.3
        st.rew_batch = torch.randn(T, N, generator=g)
        st.done_batch = (torch.rand(T, N, generator=g) < 0.1).float()
        st.compute_estimates(0.99, 0.95, True, True)
        for k in ("obs_batch", "act_batch", "log_prob_act_batch", "value_batch", "rew_batch", "done_batch",
                  "return_batch", "adv_batch"):
            out[f"opt_{tag}_{k}"] = getattr(st, k).numpy().copy()
        agent = ppo_mod.PPO(None, pol, _DummyLogger(), st, "cpu", 1, n_steps=T, n_envs=N, epoch=2, n_minibatch=4,
                            mini_batch_size=64, gamma=0.99, lmbda=0.95, learning_rate=5e-3, grad_clip_norm=0.5,
                            eps_clip=0.2, value_coef=0.5, entropy_coef=0.02, x_entropy_coef=x_coef)
        torch.manual_seed(4321)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            summary = agent.optimize()
        for n, p in pol.named_parameters():
            out[f"opt_{tag}_final/{n}"] = p.detach().numpy().copy()
        out[f"opt_{tag}_summary_keys"] = np.array(list(summary.keys()))
        out[f"opt_{tag}_summary_vals"] = np.array([float(v) for v in summary.values()])
        out[f"opt_{tag}_param_names"] = np.array(names)
        sd = agent.optimizer.state_dict()
        out[f"opt_{tag}_adam_step"] = np.array(float(sd["state"][0]["step"]))
    np.savez_compressed(os.path.join(OUT, "ppo.npz"), **out)
    print("ppo fixtures ok")


def logger_batches(seed=9, T=48, N=56, iters=5, p_done=0.06, integer=False):
    """Deterministic (reward, done) batches for the logger fixture / tests (train and validation streams)."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(iters):
        if integer:
            rew = rng.integers(-1, 3, (2, T, N)).astype(np.float32)
        else:
            rew = rng.normal(size=(2, T, N)).astype(np.float32)
        done = (rng.random((2, T, N)) < p_done).astype(np.float32)
        done[:, :, 0] = (rng.random((2, T)) < 0.3)             # an env with many short episodes
        done[:, :, N - 1] = 0                                  # an env that never finishes
        out.append((rew[0], done[0], rew[1], done[1]))
    return out


def mint_logger():
    """Rows of log-append.csv written by the reference Logger (common/logger.py) for fixed batches."""
    import csv
    import tempfile
    import contextlib
    import io
    logger_mod = ref_shim.load("common.logger")
    out = {}
    for tag, integer, max_steps in (("float", False, 7), ("int", True, 5)):
        with tempfile.TemporaryDirectory() as d:
            lg = logger_mod.Logger(56, d)
            lg.max_steps = max_steps
            for i, (r, dn, rv, dv) in enumerate(logger_batches(integer=integer)):
                lg.feed(r, dn, np.nan, rv, dv, np.nan)
                summary = {k: 0.1 * (i + 1) * (j + 1) for j, k in enumerate(
                    ["Loss/pi", "Loss/v", "Loss/entropy", "Loss/x_entropy", "Loss/atn_entropy", "Loss/atn_entropy2",
                     "Loss/sparsity", "Loss/feature_sparsity", "Loss/total"])}
                with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
                    warnings.simplefilter("ignore")
                    lg.dump(summary, 1e-3 / (i + 1))
            rows = list(csv.reader(open(os.path.join(d, "log-append.csv"))))
        out[f"{tag}_columns"] = np.array(rows[0])
        out[f"{tag}_rows"] = np.array([[float(x) if x != "" else np.nan for x in r] for r in rows[1:]])
        out[f"{tag}_num_episodes"] = np.array(lg.num_episodes)
        out[f"{tag}_reward_buffer"] = np.array(lg.episode_reward_buffer, dtype=np.float64)
        out[f"{tag}_len_buffer"] = np.array(lg.episode_len_buffer)
        out[f"{tag}_timeout_buffer"] = np.array(lg.episode_timeout_buffer)
    np.savez_compressed(os.path.join(OUT, "logger.npz"), **out)
    print("logger fixture ok:", out["float_rows"].shape, int(out["float_num_episodes"]))


CONFIG_SETS = ("boxworld-impala", "cartpole", "hard-500", "acrobot", "mountain_car", "cartpole_swing")


def mint_config():
    """The YAML sets the tests / bench run, extracted (values unchanged) from the reference's
    hyperparams/procgen/config.yml: the GPU box has no /root/reference."""
    import yaml
    with open(os.path.join(ref_shim.REFERENCE_ROOT, "hyperparams/procgen/config.yml")) as f:
        full = yaml.safe_load(f)
    sub = {k: full[k] for k in CONFIG_SETS if k in full}
    with open(os.path.join(OUT, "config_subset.yml"), "w") as f:
        f.write("# extracted by oracle/mint_golden.py from the reference's hyperparams/procgen/config.yml (values unchanged)\n")
        yaml.safe_dump(sub, f, sort_keys=False)
    print("config sets:", list(sub))


def mint_recurrent():
    """Row N4: the reference's recurrent policy (CategoricalPolicy(recurrent=True)) -- a chain of predict-time forward
    passes with done masks, and PPO.optimize on env-permuting minibatches (which does not call the GRU)."""
    storage_mod = ref_shim.load("common.storage")
    model_mod = ref_shim.load("common.model")
    policy_mod = ref_shim.load("common.policy")
    ppo_mod = ref_shim.load("agents.ppo")
    out = {}
    T, N, A, D = 16, 16, 3, 64
    torch.manual_seed(78)
    pol = policy_mod.CategoricalPolicy(model_mod.MLPModel(in_channels=9, depth=4, mid_weight=64, latent_size=D), True, A)
    names = [n for n, _ in pol.named_parameters()]
    out["param_names"] = np.array(names)
    for n, p in pol.named_parameters():
        out[f"init/{n}"] = p.detach().numpy().copy()
    g = torch.Generator().manual_seed(6)
    # --- predict chain: obs_t, done_{t-1} -> logits, value, hidden_{t+1} (agents/ppo.py:72-81) ---
    steps = 6
    obs = torch.randn(steps, N, 9, generator=g)
    done = (torch.rand(steps, N, generator=g) < 0.3).float()
    done[0] = 0
    h = torch.zeros(N, D)
    logits, values, hiddens = [], [], [h.numpy().copy()]
    with torch.no_grad():
        for t in range(steps):
            dist, v, h = pol(obs[t], h, 1 - done[t])
            logits.append(dist.logits.numpy().copy())
            values.append(v.numpy().copy())
            hiddens.append(h.numpy().copy())
    out["chain_obs"], out["chain_done_prev"] = obs.numpy(), done.numpy()
    out["chain_logits"], out["chain_value"], out["chain_hidden"] = np.stack(logits), np.stack(values), np.stack(hiddens)
    # --- optimize() ---
    st = storage_mod.Storage((9,), D, T, N, "cpu")
    st.obs_batch = torch.randn(T + 1, N, 9, generator=g)
    st.hidden_states_batch = torch.randn(T + 1, N, D, generator=g)
    st.act_batch = torch.randint(0, A, (T, N), generator=g).float()
    st.log_prob_act_batch = -torch.rand(T, N, generator=g) * 1.5 - 0.4
    st.value_batch = torch.randn(T + 1, N, generator=g) * 0.3
    st.rew_batch = torch.randn(T, N, generator=g)
    st.done_batch = (torch.rand(T, N, generator=g) < 0.1).float()
    st.compute_estimates(0.99, 0.95, True, True)
    for k in ("obs_batch", "act_batch", "log_prob_act_batch", "value_batch", "rew_batch", "done_batch", "return_batch",
              "adv_batch"):
        out[f"opt_{k}"] = getattr(st, k).numpy().copy()
    agent = ppo_mod.PPO(None, pol, _DummyLogger(), st, "cpu", 1, n_steps=T, n_envs=N, epoch=2, n_minibatch=4,
                        mini_batch_size=64, gamma=0.99, lmbda=0.95, learning_rate=5e-3, grad_clip_norm=0.5,
                        eps_clip=0.2, value_coef=0.5, entropy_coef=0.02, x_entropy_coef=0.0)
    torch.manual_seed(4321)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        summary = agent.optimize()
    for n, p in pol.named_parameters():
        out[f"final/{n}"] = p.detach().numpy().copy()
    out["summary_keys"] = np.array(list(summary.keys()))
    out["summary_vals"] = np.array([float(v) for v in summary.values()])
    sd = agent.optimizer.state_dict()
    out["adam_step"] = np.array(float(sd["state"][0]["step"]))
    out["adam_state_params"] = np.array(sorted(sd["state"].keys()))       # the GRU's four tensors never get a state
    np.savez_compressed(os.path.join(OUT, "recurrent.npz"), **out)
    print("recurrent fixtures ok; GRU unchanged by optimize:",
          all(np.array_equal(out[f"init/{n}"], out[f"final/{n}"]) for n in names if n.startswith("gru.")))


def main():
    os.makedirs(OUT, exist_ok=True)
    warnings.filterwarnings("ignore", category=SyntaxWarning)
    if len(sys.argv) > 1 and sys.argv[1] == "recurrent":
        mint_recurrent()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "logger":
        mint_logger()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "config":
        mint_config()
        return
    for fam in FAMILIES:
        mint_prevec(fam)
    mint_boxworld()
    mint_ppo()
    mint_logger()
    mint_config()
    mint_recurrent()
    print("written to", OUT, [f for f in os.listdir(OUT)])


if __name__ == "__main__":
    main()
