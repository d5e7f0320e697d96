"""Device-resident pre-vectorised env base class.

Mirrors the contract of the reference's ``PreVecEnv`` (discrete_env/pre_vec_env.py:21-219): same ctor
kwargs per family, ``reset() -> obs``, ``step(action) -> (obs, reward, done, info)`` with POST-reset obs and
PRE-reset done, the same attributes (``n_envs, n_actions, max_steps, state, n_steps, terminated, reward,
observation_space, action_space, customizable_params``) and the ``create_*`` constructors fed from the YAML sets.

What differs by design: state lives in HBM as feature-major fp32 columns, one fused CUDA kernel
(``tpp_env_step``, csrc/env_prevec.cu) does transition + truncation + auto-reset + emit, and start states come
from a counter-based Philox stream instead of numpy's PCG64 (the PCG64 bit-stream cannot be
reproduced per env; parity tests inject the reference's reset rows through ``step(..., reset_rows=...)``).
Returned tensors are CUDA tensors; ``numpy_compat=True`` returns float64 numpy arrays like the reference.
"""
from __future__ import annotations

import ctypes as C
import inspect
import math

import numpy as np
import torch

from .. import _lib
from .helper_pre_vec import assign_env_vars


class Box:
    """Minimal stand-in for gymnasium.spaces.Box (only what the reference's callers read)."""

    def __init__(self, low, high, dtype=np.float32):
        self.low = np.asarray(low, dtype=dtype)
        self.high = np.asarray(high, dtype=dtype)
        self.shape = self.low.shape
        self.dtype = np.dtype(dtype)


class Discrete:
    def __init__(self, n):
        self.n = int(n)
        self.shape = ()
        self.dtype = np.dtype(np.int64)


def _round_up(x, m):
    return (x + m - 1) // m * m


class PreVecEnv:
    """Base class; subclasses set ``family``, ``n_state``, ``n_obs``, ``low``, ``high``, ``start_low``,
    ``start_high``, ``kernel_params`` before calling ``super().__init__``."""

    family = None
    n_state = 0
    n_obs = 0
    drop_same = False
    envs_per_thread = 0     # kernel hint (tpp_env_cfg.p[7]): 0 = default (4, 128-bit accesses); ALU-bound families use 1

    def __init__(self, n_envs, n_actions, env_name, max_steps=500, seed=0, render_mode=None, device="cuda",
                 numpy_compat=False):
        if n_envs < 2:
            raise Exception("n_envs must be greater than or equal to 2")   # pre_vec_env.py:41-42
        if render_mode is not None:
            raise NotImplementedError("rendering is outside the hot path (SURVEY section 2, row 24)")
        self.env_name = env_name
        self.n_envs = int(n_envs)
        self.n_actions = int(n_actions)
        self.max_steps = int(max_steps)
        self.render_mode = None
        self.numpy_compat = numpy_compat
        self.device = torch.device(device)
        self.np_random_seed = seed
        self.non_drop_index = self.high != self.low
        if not self.drop_same:
            self.non_drop_index = self.high == self.high
        self.action_space = Discrete(self.n_actions)
        self.observation_space = Box(self.low[self.non_drop_index], self.high[self.non_drop_index])
        self.n_inputs = self.n_state

        self.ld = _round_up(self.n_envs, 32)
        dev = self.device
        # two private rollout slots (ping-pong) used when the env is stepped stand-alone; inside PPO.train the
        # kernel writes straight into the Storage slots instead (step_into).
        self._slots = torch.zeros(2, self.n_obs, self.ld, dtype=torch.float32, device=dev)
        self._cur = 0
        self._dyn = torch.zeros(4, self.ld, dtype=torch.float32, device=dev) if self.family == "acrobot" else None
        self._step_ctr = torch.zeros(self.ld, dtype=torch.int32, device=dev)
        self._rew = torch.zeros(self.ld, dtype=torch.float32, device=dev)
        self._done = torch.zeros(self.ld, dtype=torch.uint8, device=dev)
        self._act = torch.zeros(self.ld, dtype=torch.int32, device=dev)
        self._tick = torch.zeros(1, dtype=torch.int64, device=dev)   # uint64 on the device side
        self._cfg = self._make_cfg(seed if seed is not None else 0)
        self._obs_view = None
        self.reset(seed=seed)

    # ------------------------------------------------------------------------------------------
    def _make_cfg(self, seed):
        cfg = _lib.EnvCfg()
        cfg.family = _lib.FAMILY[self.family]
        cfg.n_envs = self.n_envs
        cfg.max_steps = self.max_steps
        cfg.n_state = self.n_state
        cfg.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        for i, (lo, hi) in enumerate(zip(self.start_low, self.start_high)):
            cfg.start_low[i] = float(lo)
            cfg.start_high[i] = float(hi)
        for i, v in enumerate(self.kernel_params):
            cfg.p[i] = float(v)
        cfg.p[7] = float(self.envs_per_thread)
        return cfg

    def _rows_to_device(self, rows):
        """[N, n_state] (numpy / tensor, any float dtype) -> feature-major fp32 [n_state, ld] on the device."""
        if rows is None:
            return None
        r = torch.as_tensor(np.asarray(rows) if not torch.is_tensor(rows) else rows)
        r = r.to(self.device, torch.float32)
        assert r.shape == (self.n_envs, self.n_state), r.shape
        out = torch.zeros(self.n_state, self.ld, dtype=torch.float32, device=self.device)
        out[:, :self.n_envs] = r.t()
        return out

    def _actions_to_device(self, action):
        if torch.is_tensor(action) and action.is_cuda:
            assert action.numel() == self.n_envs, \
                f"number of actions ({action.numel()}) must match n_envs ({self.n_envs})"
            self._act[:self.n_envs] = action.reshape(-1).to(torch.int32)
            return self._act
        a = np.asarray(action.cpu() if torch.is_tensor(action) else action)
        assert a.size == self.n_envs, f"number of actions ({a.size}) must match n_envs ({self.n_envs})"
        assert np.all(a < self.n_actions), f"action must be less than n_actions ({self.n_actions})"
        self._act[:self.n_envs] = torch.from_numpy(a.reshape(-1).astype(np.int32)).to(self.device)
        return self._act

    # ------------------------------------------------------------------------------------------
    # Fast path used by the fused rollout: all pointers are rollout-slot rows on the device.
    # ------------------------------------------------------------------------------------------
    def step_into(self, obs_in, obs_out, action_i32, rew_out, done_out, t_offset=0, reset_rows=None, ld=None):
        _lib.call("tpp_env_step", C.byref(self._cfg), _lib.ptr(obs_in), _lib.ptr(obs_out), _lib.ptr(self._dyn),
                  _lib.ptr(action_i32), _lib.ptr(self._step_ctr), _lib.ptr(rew_out), _lib.ptr(done_out),
                  _lib.ptr(reset_rows), _lib.ptr(self._tick), int(t_offset), int(ld or self.ld), _lib.stream_ptr())

    def reset_into(self, obs_out, t_offset=0, reset_rows=None, ld=None):
        _lib.call("tpp_env_reset", C.byref(self._cfg), _lib.ptr(obs_out), _lib.ptr(self._dyn),
                  _lib.ptr(self._step_ctr), _lib.ptr(reset_rows), _lib.ptr(self._tick), int(t_offset),
                  int(ld or self.ld), _lib.stream_ptr())

    def reset_rollout(self, storage):
        """Reset all envs straight into slot 0 of a Storage's rollout (start of PPO.train)."""
        assert storage.ld == self.ld and not storage.is_image and storage.obs_width == self.n_obs
        self.reset_into(storage.obs_slot(0))
        self.advance_tick()

    def rollout_step(self, storage, t):
        """One fused env step reading slot t / writing slot t+1, reward[t], done[t] of the rollout."""
        self.step_into(storage.obs_slot(t), storage.obs_slot(t + 1), storage.act_i32[t], storage.rew[t],
                       storage.done_u8[t], t_offset=t)

    def advance_tick(self, delta=1):
        _lib.call("tpp_tick_advance", _lib.ptr(self._tick), int(delta), _lib.stream_ptr())

    # ------------------------------------------------------------------------------------------
    # Reference API
    # ------------------------------------------------------------------------------------------
    def seed(self, seed=None):
        self.np_random_seed = seed
        if seed is not None:
            self._cfg.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
            self._tick.zero_()
        return [seed]

    def reset(self, *, seed=None, options=None, reset_rows=None):
        self.seed(seed)
        self.reset_into(self._slots[self._cur], reset_rows=self._rows_to_device(reset_rows))
        self.advance_tick()
        self._done.fill_(1)
        return self._get_ob()

    def step(self, action, reset_rows=None):
        act = self._actions_to_device(action)
        nxt = self._cur ^ 1
        self.step_into(self._slots[self._cur], self._slots[nxt], act, self._rew, self._done,
                       reset_rows=self._rows_to_device(reset_rows))
        self.advance_tick()
        self._cur = nxt
        rew, done = self.reward, self.terminated
        return self._get_ob(), rew, done, self.info

    def _get_ob(self):
        ob = self._slots[self._cur][:, :self.n_envs].t()
        if self.drop_same:
            ob = ob[:, torch.as_tensor(self.non_drop_index, device=self.device)]
        if self.numpy_compat:
            return ob.cpu().numpy().astype(np.float64)
        return ob

    @property
    def state(self):
        """[N, n_state] like the reference's ``self.state`` (a fresh tensor, not an alias)."""
        cur = self._slots[self._cur][:, :self.n_envs]
        if self._dyn is not None:   # acrobot: 4 dynamic columns + the parameter columns of the observation
            st = torch.cat((self._dyn[:, :self.n_envs], cur[self.n_obs - (self.n_state - 4):]), 0).t()
        else:
            st = cur.t()
        return st.cpu().numpy().astype(np.float64) if self.numpy_compat else st

    @property
    def n_steps(self):
        v = self._step_ctr[:self.n_envs]
        return v.cpu().numpy().astype(np.float64) if self.numpy_compat else v

    @property
    def terminated(self):
        v = self._done[:self.n_envs].bool()
        return v.cpu().numpy() if self.numpy_compat else v

    @property
    def reward(self):
        v = self._rew[:self.n_envs]
        return v.cpu().numpy().astype(np.float64) if self.numpy_compat else v

    @property
    def info(self):
        if self.numpy_compat:   # the reference's per-env dict list (O(N) python: compat mode only)
            r = self._rew[:self.n_envs].cpu().numpy().astype(np.float64)
            return [{"env_reward": r[i]} for i in range(self.n_envs)]
        return EnvInfo(self._rew[:self.n_envs])

    def get_info(self):
        return self.info

    def get_params(self, suffix=""):
        return {f"{name}{suffix}": getattr(self, name) for name in self.customizable_params}

    def save(self):
        np.save(f"{self.env_name}.npy", np.asarray(self.state.cpu() if torch.is_tensor(self.state) else self.state))

    def render(self):
        raise NotImplementedError("rendering is outside the hot path")

    def close(self):
        pass

    def get_action_lookup(self):
        raise NotImplementedError

    def get_ob_names(self):
        raise NotImplementedError


class EnvInfo:
    """Lazy replacement of the reference's ``[{'env_reward': r_i} ...]`` list: indexable, device-backed."""

    def __init__(self, env_reward):
        self.env_reward = env_reward

    def __len__(self):
        return self.env_reward.numel()

    def __getitem__(self, i):
        return {"env_reward": self.env_reward[i]}


def create_pre_vec(args, hyperparameters, param_range, env_cons, is_valid):
    """discrete_env/pre_vec_env.py:204-214: YAML set -> ctor kwargs by signature filtering."""
    seed = getattr(args, "seed", 0) if args is not None else 0
    n_envs = hyperparameters.get("n_envs", 32)
    env_args = assign_env_vars(hyperparameters, is_valid, param_range)
    env_args = filter_out_non_relevant_params(env_args, env_cons)
    env_args["n_envs"] = n_envs
    env_args["seed"] = seed
    for k in ("device", "numpy_compat"):
        if k in hyperparameters:
            env_args[k] = hyperparameters[k]
    return env_cons(**env_args)


def filter_out_non_relevant_params(env_args, env_cons):
    params = inspect.signature(env_cons.__init__).parameters.keys()
    return {k: v for k, v in env_args.items() if k in params}
