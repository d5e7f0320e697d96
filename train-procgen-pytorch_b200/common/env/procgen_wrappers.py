"""Host-stepped (Procgen-style) envs behind the GPU rollout: the reference's wrapper stack
VecExtractDictObs -> VecNormalize(ob=False) -> TransposeFrame -> ScaledFloatFrame [-> ActionWrapper]
(common/env/procgen_wrappers.py:265-446, create_procgen_env :549-607) folded into ONE object that keeps the closed C
engine on the host and everything else on the device:

* frames stay uint8 NHWC from the engine to the rollout slot (double-buffered pinned H2D, ``Storage.stage_obs``);
  TransposeFrame + ScaledFloatFrame happen where the policy reads them (``tpp_frames_to_obs`` / ``tpp_gather_img``);
* ``VecNormalize``'s running return statistics live on the device and are applied to the T steps of a rollout after its
  last step (``finish_rollout``: the per-step ``tpp_vecnormalize_step`` kernel, same float64 recurrences) -- the
  normalised reward never feeds back into the rollout, so it does not sit on the step's critical path;
* ``ActionWrapper`` (de-duplicated action names) is a lookup applied to the N sampled actions after their D2H copy.

The reference VecEnv API (``reset`` / ``step`` returning float NCHW observations in [0, 1], normalised rewards and
``info[i]['env_reward']``) is kept for callers that drive the env themselves (numpy, host arithmetic as upstream).
"""
from __future__ import annotations

import numpy as np
import torch

from ... import _lib
from ...discrete_env.pre_vec_env import Box, Discrete


def match(a, b, dtype=np.int32):
    """helper_local.py:65-70: index in ``b`` of every element of ``a`` that occurs in ``b``."""
    a, b = list(np.asarray(a).tolist()), list(np.asarray(b).tolist())
    return np.array([b.index(x) for x in a if x in b], dtype=dtype)


def unique_action_mapping(action_names):
    """ActionWrapper.__init__ (common/env/procgen_wrappers.py:427-436): the sorted unique action names and, for each of
    them, the first engine action that carries the name."""
    names = np.asarray(action_names)
    unique = np.unique(names)
    return unique, match(unique, names)


class RunningMeanStd:
    """common/env/procgen_wrappers.py:282-311 (host copy used by the VecEnv API path)."""

    def __init__(self, epsilon=1e-4):
        self.mean, self.var, self.count = 0.0, 1.0, epsilon

    def update(self, x):
        bm, bv, bc = np.mean(x, axis=0), np.var(x, axis=0), x.shape[0]
        delta, tot = bm - self.mean, self.count + bc
        m2 = self.var * self.count + bv * bc + np.square(delta) * self.count * bc / tot
        self.mean, self.var, self.count = self.mean + delta * bc / tot, m2 / tot, tot


class StagedVecEnv:
    """``venv``: the host engine after VecExtractDictObs -- ``reset() -> frames`` and ``step(actions) -> (frames, rew,
    done, infos)`` with uint8 NHWC frames (a dict observation with an ``'rgb'`` entry is unwrapped).  ``action_names``
    + ``reduce_duplicate_actions=True`` reproduce ActionWrapper."""

    stages_raw_frames = True

    def __init__(self, venv, n_envs=None, normalize_rew=True, gamma=0.99, cliprew=10.0, epsilon=1e-8, action_names=None,
                 reduce_duplicate_actions=False, device="cuda"):
        self.venv = venv
        self.num_envs = int(n_envs or getattr(venv, "num_envs", None) or getattr(venv, "n"))
        self.normalize_rew, self.gamma, self.cliprew, self.epsilon = normalize_rew, gamma, cliprew, epsilon
        shape = tuple(venv.observation_space.shape)
        h, w, c = shape if shape[-1] in (1, 3, 4) else (shape[1], shape[2], shape[0])
        self.observation_space = Box(np.zeros((c, h, w)), np.ones((c, h, w)), dtype=np.float32)
        n_act = int(venv.action_space.n)
        self.action_mapping = None
        if reduce_duplicate_actions:
            assert action_names is not None and len(action_names) == n_act
            self.unique_actions, self.action_mapping = unique_action_mapping(action_names)
            n_act = len(self.unique_actions)
        self.action_space = Discrete(n_act)
        self.device = torch.device(device)
        # device VecNormalize state (rollout path) and its host twin (VecEnv API path)
        self._ret = torch.zeros(self.num_envs, dtype=torch.float64, device=self.device)
        self._rms = torch.tensor([0.0, 1.0, 1e-4], dtype=torch.float64, device=self.device)
        self.ret, self.ret_rms = np.zeros(self.num_envs), RunningMeanStd()

    # ---- host engine access ------------------------------------------------------------------------------------
    @staticmethod
    def _rgb(obs):
        return obs["rgb"] if isinstance(obs, dict) else obs

    def map_actions(self, actions):
        a = np.asarray(actions)
        return self.action_mapping[a] if self.action_mapping is not None else a

    def host_reset(self):
        return np.ascontiguousarray(self._rgb(self.venv.reset()))

    def host_step(self, actions):
        """(uint8 NHWC frames, RAW rewards, dones, infos) -- what the device rollout stages."""
        obs, rew, done, info = self.venv.step(self.map_actions(actions))
        return np.ascontiguousarray(self._rgb(obs)), np.asarray(rew, dtype=np.float32), np.asarray(done), info

    def finish_rollout(self, storage):
        """VecNormalize over the rollout's T steps on the device: raw rewards ``storage.env_rew`` -> ``storage.rew``."""
        if not self.normalize_rew:
            storage.rew.copy_(storage.env_rew)
            return
        for t in range(storage.num_steps):
            _lib.call("tpp_vecnormalize_step", _lib.ptr(self._ret), _lib.ptr(self._rms), _lib.ptr(storage.env_rew[t]),
                      0, _lib.ptr(storage.done_u8[t]), _lib.ptr(storage.rew[t]), self.num_envs, self.gamma,
                      self.cliprew, self.epsilon, _lib.stream_ptr())
        storage.n_launches += storage.num_steps

    # ---- reference VecEnv API (host arithmetic, as upstream) -----------------------------------------------------
    def _obs(self, frames):
        return frames.transpose(0, 3, 1, 2) / 255.0          # TransposeFrame + ScaledFloatFrame (:398-419)

    def reset(self):
        self.ret = np.zeros(self.num_envs)
        return self._obs(self.host_reset())

    def step(self, actions):
        frames, rew, done, info = self.host_step(actions)
        rew = rew.astype(np.float64)
        if isinstance(info, (list, tuple)):
            for i in range(len(info)):
                if isinstance(info[i], dict):
                    info[i]["env_reward"] = rew[i]
        self.ret = self.ret * self.gamma + rew                  # VecNormalize.step_wait (:332-342)
        if self.normalize_rew:
            self.ret_rms.update(self.ret)
            rew = np.clip(rew / np.sqrt(self.ret_rms.var + self.epsilon), -self.cliprew, self.cliprew)
        self.ret[np.asarray(done, dtype=bool)] = 0.0
        return self._obs(frames), rew, done, info

    def close(self):
        if hasattr(self.venv, "close"):
            self.venv.close()
