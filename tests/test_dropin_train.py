"""Drop-in surfaces around PPO.train (SURVEY 8b "Config", 8f N3 / N4):

* the reference's YAML sets are consumed UNCHANGED (helper_local.get_hyperparams / train_ppo; fixture
  tests/golden/config_subset.yml is extracted from hyperparams/procgen/config.yml by oracle/mint_golden.py);
* validation-env rollouts (agents/ppo.py:241-252) fill the ``val_*`` columns of log-append.csv;
* checkpoints (agents/ppo.py:271-276, train.py:257-263): ``model_<t>.pth`` round-trips through the reference-shaped
  torch modules and ``torch.optim.Adam`` and resumes to a bit-identical next ``optimize()``;
* host-stepped envs through ``StagedVecEnv`` (common/env/procgen_wrappers.py:314-355,391-446): raw rewards reach the
  logger, VecNormalize statistics equal the oracle's, ActionWrapper's mapping is applied."""
import csv
import glob
import os

import numpy as np
import pytest
import torch

from oracle import ref_shim

GOLDEN_CFG = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "config_subset.yml")


def test_yaml_sets_load_unchanged():
    from tpp_b200.helper_local import get_hyperparams
    hp = get_hyperparams("cartpole", GOLDEN_CFG)
    assert hp["n_envs"] == 256 and hp["n_minibatch"] == 16 and hp["architecture"] == "mlpmodel" and hp["depth"] == 4
    hp = get_hyperparams("boxworld-impala", GOLDEN_CFG)
    assert hp["grid_size"] == 12 and hp["goal_length_v"] == 5 and hp["architecture"] == "impala"
    assert get_hyperparams("hard-500", GOLDEN_CFG)["mini_batch_per_epoch"] == 8


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present (GPU box)")
def test_yaml_fixture_equals_reference_file():
    import yaml
    from tpp_b200.helper_local import get_hyperparams
    ref_path = os.path.join(ref_shim.REFERENCE_ROOT, "hyperparams/procgen/config.yml")
    sub = yaml.safe_load(open(GOLDEN_CFG))
    assert len(sub) >= 4
    for name, want in sub.items():
        assert get_hyperparams(name, ref_path) == want


def test_action_wrapper_mapping_matches_reference_semantics():
    """ActionWrapper (common/env/procgen_wrappers.py:427-436): sorted unique names -> first engine action of that name."""
    from tpp_b200.common.env.procgen_wrappers import unique_action_mapping
    names = np.array(["LEFT_DOWN", "LEFT", "LEFT_UP", "DOWN", "", "UP", "RIGHT_DOWN", "RIGHT", "RIGHT_UP", "RIGHT",
                      "LEFT", "UP", "DOWN", "LEFT_UP", "RIGHT_UP"])        # helper_local.get_action_names on Procgen
    uniq, mapping = unique_action_mapping(names)
    assert list(uniq) == sorted(set(names.tolist()))
    assert all(names[m] == u for m, u in zip(mapping, uniq))
    assert all(m == names.tolist().index(u) for m, u in zip(mapping, uniq))


def _read_csv(logdir):
    rows = list(csv.reader(open(os.path.join(logdir, "log-append.csv"))))
    return rows[0], np.array([[float(x) if x not in ("", "nan") else np.nan for x in r] for r in rows[1:]])


@pytest.mark.gpu
def test_cartpole_yaml_set_trains_with_validation_env_and_checkpoints(tmp_path):
    """BASELINE configs[0] exactly as the YAML states it (256 envs x 256 steps, 16 minibatches), validation env on."""
    from tpp_b200.helper_local import Args, get_hyperparams, train_ppo
    hp = get_hyperparams("cartpole", GOLDEN_CFG)
    iters = 4
    args = Args(seed=6033, use_valid_env=True, num_checkpoints=2, logdir=str(tmp_path),
                num_timesteps=hp["n_envs"] * hp["n_steps"] * iters)
    agent = train_ppo(args, hp, "cartpole")
    assert agent.t == args.num_timesteps and agent.n_minibatch == 16 and agent.storage_valid is not None
    cols, rows = _read_csv(str(tmp_path))
    assert rows.shape == (iters, len(cols)) and cols[:3] == ["timesteps", "wall_time", "num_episodes"]
    c = {k: i for i, k in enumerate(cols)}
    assert (rows[:, c["timesteps"]] == np.arange(1, iters + 1) * 65536).all()
    # training and validation rollouts both produced episodes (cartpole reward == 1 per step -> return == length)
    for pre in ("", "val_"):
        assert np.isfinite(rows[-1, c[pre + "mean_episode_rewards"]])
        assert rows[-1, c[pre + "mean_episode_rewards"]] == rows[-1, c[pre + "mean_episode_len"]]
        # (the reference takes np.min / np.max with initial=0, common/logger.py:180-186: the minimum column is 0)
        assert rows[-1, c[pre + "min_episode_len"]] == 0 and 1 <= rows[-1, c[pre + "max_episode_len"]] <= 500
    # the validation env draws its physics from the `_v` parameter ranges (discrete_env/helper_pre_vec.py:49-62)
    assert agent.env_valid.degrees == hp["degrees_v"] and agent.env.degrees == 12
    # linear lr decay applied after every iteration (agents/ppo.py:267): column = lr of the NEXT update
    np.testing.assert_allclose(rows[:, c["learning_rate"]], hp["learning_rate"] * (1 - np.arange(1, iters + 1) / iters))
    # reference rule (agents/ppo.py:212-215,271-276): saved when t EXCEEDS the k-th of num_checkpoints evenly spaced
    # marks, so the mark at num_timesteps itself never fires
    ck = sorted(os.path.basename(f) for f in glob.glob(str(tmp_path / "model_*.pth")))
    assert ck == ["model_196608.pth"]


@pytest.mark.gpu
def test_checkpoint_roundtrip_resumes_the_run(tmp_path):
    from oracle import ppo as oppo
    from tpp_b200.helper_local import Args, get_hyperparams, train_ppo
    hp = dict(get_hyperparams("cartpole", GOLDEN_CFG), n_envs=64, n_steps=32, n_minibatch=2, mini_batch_size=512)
    it = 64 * 32
    args = Args(seed=3, num_checkpoints=3, logdir=str(tmp_path), num_timesteps=it * 3)
    a = train_ppo(args, hp, "cartpole", train=False)
    rng_states, rollouts, inner = [], [], a.optimize
    NAMES = ("obs_fm", "act_i32", "logp", "rew", "done_u8", "value")

    def recording_optimize(*p, **k):           # what every optimize() consumed: generator state + rollout
        rng_states.append(torch.get_rng_state())
        rollouts.append({n: getattr(a.storage, n).clone() for n in NAMES})
        return inner(*p, **k)
    a.optimize = recording_optimize
    a.train(args.num_timesteps)
    # marks at 1, 2, 3 iterations; a checkpoint is written when t exceeds a mark: after iterations 2 and 3
    assert sorted(os.path.basename(f) for f in glob.glob(str(tmp_path / "model_*.pth"))) == \
        [f"model_{2 * it}.pth", f"model_{3 * it}.pth"]
    path = str(tmp_path / f"model_{2 * it}.pth")
    ck = torch.load(path, map_location="cpu")
    assert set(ck) == {"model_state_dict", "optimizer_state_dict"}
    # (1) the reference-shaped torch modules + torch.optim.Adam accept the checkpoint as is (train.py:257-263)
    ref = oppo.OraclePolicy(oppo.OracleMLP(9, hp["depth"], hp["mid_weight"], hp["latent_size"]), 2)
    ref.load_state_dict(ck["model_state_dict"])
    opt = oppo.make_adam(ref, hp["learning_rate"])
    opt.load_state_dict(ck["optimizer_state_dict"])
    st0 = opt.state_dict()["state"][0]
    assert float(st0["step"]) == 2 * hp["epoch"] * 2 and st0["exp_avg"].shape == ref.embedder.model[0].weight.shape
    # written after adjust_lr like the reference's (agents/ppo.py:267-276): the rate of the NEXT update
    np.testing.assert_allclose(ck["optimizer_state_dict"]["param_groups"][0]["lr"],
                               hp["learning_rate"] * (1 - 2 / 3), rtol=1e-12)
    # (2) a fresh agent that loads it and is handed iteration 3's rollout + generator state reproduces iteration 3 of
    # the original run (gradients meet in fp32 atomics, so equality is to rounding, not bitwise)
    def resumed(load_optimizer):
        b = train_ppo(Args(seed=3, model_file=path if load_optimizer else None), hp, "cartpole", train=False)
        if not load_optimizer:
            b.policy.load_state_dict(ck["model_state_dict"])
            b.optimizer.param_groups[0]["lr"] = ck["optimizer_state_dict"]["param_groups"][0]["lr"]
        for name in NAMES:
            getattr(b.storage, name).copy_(rollouts[2][name])
        b.storage.compute_estimates(b.gamma, b.lmbda, True, True)
        torch.set_rng_state(rng_states[2])
        b.optimize()
        return b
    b = resumed(True)
    assert b.optimizer.step_count == 3 * hp["epoch"] * 2
    want = a.policy.flat.cpu().numpy()
    np.testing.assert_allclose(b.policy.flat.cpu().numpy(), want, rtol=2e-5, atol=2e-7)
    # ... which needs the Adam moments and step count: weights alone do not get there
    c = resumed(False)
    assert np.abs(c.policy.flat.cpu().numpy() - want).max() > 50 * np.abs(b.policy.flat.cpu().numpy() - want).max()


@pytest.mark.gpu
def test_boxworld_impala_yaml_set_runs_unchanged(tmp_path):
    """The `boxworld-impala` set (config.yml:575-601) as written: 256 envs, grid 12, IMPALA policy on 14x14 frames."""
    from tpp_b200.common.engine import ImpalaEngineTC
    from tpp_b200.helper_local import Args, get_hyperparams, train_ppo
    hp = get_hyperparams("boxworld-impala", GOLDEN_CFG)
    args = Args(seed=6033, num_levels=500, use_valid_env=True, logdir=str(tmp_path),
                num_timesteps=hp["n_envs"] * hp["n_steps"] * 2)
    agent = train_ppo(args, hp, "boxworld")
    assert isinstance(agent.engine, ImpalaEngineTC) and agent.env.env.n == 12 and agent.env.env.n_levels == 500
    assert agent.env_valid.env.n_levels == 0 and agent.env_valid.env.start_seed == 501     # create_box_world.py:103-124
    cols, rows = _read_csv(str(tmp_path))
    c = {k: i for i, k in enumerate(cols)}
    assert rows.shape[0] == 2 and np.isfinite(rows[:, c["loss_total"]]).all()
    assert np.isfinite(rows[:, c["loss_feature_sparsity"]]).all()          # IMPALA logs it, MLP leaves NaN


class FakeProcgenRaw:
    """Host engine stand-in with Procgen's contract after VecExtractDictObs: uint8 NHWC frames, float rewards, 15
    engine actions of which several share a name (ActionWrapper's case)."""
    NAMES = ["LEFT_DOWN", "LEFT", "LEFT_UP", "DOWN", "", "UP", "RIGHT_DOWN", "RIGHT", "RIGHT_UP", "RIGHT", "LEFT", "UP",
             "DOWN", "LEFT_UP", "RIGHT_UP"]

    def __init__(self, n, hw=(64, 64), seed=0):
        from tpp_b200.discrete_env.pre_vec_env import Box, Discrete
        self.num_envs, self.hw = n, hw
        self.observation_space = Box(np.zeros((*hw, 3)), np.full((*hw, 3), 255), dtype=np.uint8)
        self.action_space = Discrete(15)
        self.rng = np.random.default_rng(seed)
        self.seen_actions, self.raw_log, self.done_log = set(), [], []

    def _frames(self):
        return self.rng.integers(0, 256, (self.num_envs, *self.hw, 3), dtype=np.uint8)

    def reset(self):
        return {"rgb": self._frames()}

    def step(self, act):
        assert act.shape == (self.num_envs,)
        self.seen_actions |= set(int(a) for a in act)
        rew = ((self.rng.random(self.num_envs) < 0.05) * 10.0).astype(np.float32)
        done = self.rng.random(self.num_envs) < 0.08
        self.raw_log.append(rew.copy()); self.done_log.append(done.copy())
        return {"rgb": self._frames()}, rew, done, [{} for _ in range(self.num_envs)]


@pytest.mark.gpu
def test_staged_host_env_pipeline_normalises_on_device_and_logs_raw_rewards(tmp_path):
    from oracle.boxworld import VecNormalizeOracle
    from tpp_b200.agents.ppo import PPO
    from tpp_b200.common.env.procgen_wrappers import StagedVecEnv
    from tpp_b200.common.logger import Logger, close_episodes
    from tpp_b200.common.model import ImpalaModel
    from tpp_b200.common.policy import CategoricalPolicy
    from tpp_b200.common.storage import Storage
    N, T, iters = 16, 8, 3
    raw_env, raw_valid = FakeProcgenRaw(N, seed=0), FakeProcgenRaw(N, seed=1)
    env = StagedVecEnv(raw_env, normalize_rew=True, gamma=0.999, action_names=FakeProcgenRaw.NAMES,
                       reduce_duplicate_actions=True)
    env_v = StagedVecEnv(raw_valid, normalize_rew=True, gamma=0.999, action_names=FakeProcgenRaw.NAMES,
                         reduce_duplicate_actions=True)
    assert env.action_space.n == 9 and env.observation_space.shape == (3, 64, 64)
    torch.manual_seed(3)
    pol = CategoricalPolicy(ImpalaModel(3), False, env.action_space.n).to("cuda").flatten_()
    st, sv = Storage((3, 64, 64), 256, T, N, "cuda"), Storage((3, 64, 64), 256, T, N, "cuda")
    lg = Logger(N, str(tmp_path))
    lg.max_steps = 1000
    agent = PPO(env, pol, lg, st, "cuda", 0, env_valid=env_v, storage_valid=sv, n_steps=T, n_envs=N, epoch=1,
                n_minibatch=2, mini_batch_size=64, gamma=0.999, learning_rate=5e-4)
    agent.train(T * N * iters)
    assert agent.t == T * N * iters and torch.isfinite(pol.flat).all()
    # ActionWrapper: only first-of-name engine actions ever reach the engine
    assert raw_env.seen_actions <= set(int(m) for m in env.action_mapping)
    # the frames of the last step were staged as uint8 and the double-buffered copies kept order: slot T == last obs
    assert st.frames.dtype == torch.uint8
    # rewards: the rollout holds VecNormalize's output (device kernel == float64 oracle restated from
    # common/env/procgen_wrappers.py:314-355), the logger saw the RAW rewards
    vn = VecNormalizeOracle(N, gamma=0.999)
    want = np.stack([vn.step(r.astype(np.float64), d) for r, d in zip(raw_env.raw_log, raw_env.done_log)])
    np.testing.assert_allclose(st.rew[:, :N].cpu().numpy(), want[-T:], rtol=1e-6, atol=1e-7)
    def fed_per_iteration(e):        # the logger is fed once per rollout: env-major inside every T-step batch
        run_r, run_l, out = np.zeros(N), np.zeros(N, dtype=np.int64), []
        for i in range(iters):
            r, _ = close_episodes(np.stack(e.raw_log[i * T:(i + 1) * T]), np.stack(e.done_log[i * T:(i + 1) * T]), run_r,
                                  run_l)
            out += list(r)
        return np.array(out)
    rets = fed_per_iteration(raw_env)
    assert lg.num_episodes == len(rets)
    np.testing.assert_allclose(np.array(lg.episode_reward_buffer), rets[-40:])
    np.testing.assert_allclose(np.array(lg.episode_reward_buffer_v), fed_per_iteration(raw_valid)[-40:])
    cols, rows = _read_csv(str(tmp_path))
    assert rows.shape[0] == iters
