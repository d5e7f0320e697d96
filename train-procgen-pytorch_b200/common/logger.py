"""Episode statistics + CSV logging with the reference Logger's interface and column schema
(common/logger.py:13-236): ``feed(rew_batch, done_batch, true_mean_reward, rew_batch_v, done_batch_v,
true_mean_reward_v)`` once per iteration, ``dump(summary, lr)``, ``episode_reward_buffer``, ``logdir``.

The reference walks the [T, N] batches with an O(T*N) Python double loop (:119-147); here the episode boundaries are
found with vectorised numpy on the arrays ``Storage.fetch_log_data`` returns (same episode accounting: a running
return per env, closed at every done flag, last 40 episodes kept).  wandb / pandas are not used on the hot path."""
from __future__ import annotations

import csv
import os
import time
from collections import deque

import numpy as np

LOSS_KEYS = ["Loss/pi", "Loss/v", "Loss/entropy", "Loss/x_entropy", "Loss/atn_entropy", "Loss/atn_entropy2",
             "Loss/sparsity", "Loss/feature_sparsity", "Loss/total"]
EPISODE_KEYS = ["max_episode_rewards", "mean_episode_rewards", "min_episode_rewards", "max_episode_len",
                "mean_episode_len", "min_episode_len", "mean_timeouts"]


class Logger:
    def __init__(self, n_envs, logdir=None, use_wandb=False, has_vq=False, transition_model=False, double_graph=False,
                 ppo_pure=False, IPL=False, sae=False):
        self.start_time = time.time()
        self.n_envs, self.logdir = n_envs, logdir
        self.episode_rewards = np.zeros(n_envs)
        self.episode_lens = np.zeros(n_envs, dtype=np.int64)
        self.episode_rewards_v = np.zeros(n_envs)
        self.episode_lens_v = np.zeros(n_envs, dtype=np.int64)
        self.episode_timeout_buffer = deque(maxlen=40)
        self.episode_len_buffer = deque(maxlen=40)
        self.episode_reward_buffer = deque(maxlen=40)
        self.episode_timeout_buffer_v = deque(maxlen=40)
        self.episode_len_buffer_v = deque(maxlen=40)
        self.episode_reward_buffer_v = deque(maxlen=40)
        self.true_mean_reward = self.true_mean_reward_v = np.nan
        self.max_steps = 10 ** 9
        self.columns = (["timesteps", "wall_time", "num_episodes"] + EPISODE_KEYS + ["val_" + k for k in EPISODE_KEYS]
                        + ["true_mean_reward", "val_true_mean_reward", "learning_rate"] + LOSS_KEYS)
        self.timesteps, self.num_episodes = 0, 0
        self.rows = []
        if logdir:
            os.makedirs(logdir, exist_ok=True)
            with open(os.path.join(logdir, "log-append.csv"), "w", newline="") as f:
                csv.writer(f).writerow(self.columns)

    @staticmethod
    def _episodes(rew, done, run_ret, run_len):
        """Close episodes at done flags.  rew, done: [T, N].  Returns (returns, lengths) in (t, env) order and
        updates the running accumulators in place — the same bookkeeping as the reference's double loop."""
        T, N = rew.shape
        csum = np.cumsum(rew, axis=0)
        rets, lens = [], []
        t_idx, e_idx = np.nonzero(done > 0)
        last_t = np.full(N, -1)
        order = np.lexsort((e_idx, t_idx))
        for t, e in zip(t_idx[order], e_idx[order]):
            start = last_t[e]
            seg = csum[t, e] - (csum[start, e] if start >= 0 else 0.0)
            rets.append(run_ret[e] + seg if start < 0 else seg)
            lens.append((run_len[e] if start < 0 else 0) + (t - start))
            last_t[e] = t
        for e in range(N):
            if last_t[e] < 0:
                run_ret[e] += csum[-1, e]
                run_len[e] += T
            else:
                run_ret[e] = csum[-1, e] - csum[last_t[e], e]
                run_len[e] = T - 1 - last_t[e]
        return rets, lens

    def feed(self, rew_batch, done_batch, true_mean_reward=np.nan, rew_batch_v=None, done_batch_v=None,
             true_mean_reward_v=np.nan):
        T, N = rew_batch.shape
        rets, lens = self._episodes(np.asarray(rew_batch, dtype=np.float64), np.asarray(done_batch),
                                    self.episode_rewards, self.episode_lens)
        for r, l in zip(rets, lens):
            self.episode_reward_buffer.append(r)
            self.episode_len_buffer.append(l)
            self.episode_timeout_buffer.append(1 if l >= self.max_steps else 0)
        self.num_episodes += len(rets)
        if rew_batch_v is not None:
            rets, lens = self._episodes(np.asarray(rew_batch_v, dtype=np.float64), np.asarray(done_batch_v),
                                        self.episode_rewards_v, self.episode_lens_v)
            for r, l in zip(rets, lens):
                self.episode_reward_buffer_v.append(r)
                self.episode_len_buffer_v.append(l)
                self.episode_timeout_buffer_v.append(1 if l >= self.max_steps else 0)
        self.true_mean_reward, self.true_mean_reward_v = true_mean_reward, true_mean_reward_v
        self.timesteps += T * N

    @staticmethod
    def _stats(rew_buf, len_buf, to_buf):
        if len(rew_buf) == 0:
            return [np.nan] * 7
        return [np.max(rew_buf), np.mean(rew_buf), np.min(rew_buf), np.max(len_buf), np.mean(len_buf), np.min(len_buf),
                np.mean(to_buf)]

    def dump(self, summary=None, lr=None):
        summary = summary or {}
        row = ([self.timesteps, time.time() - self.start_time, self.num_episodes]
               + self._stats(self.episode_reward_buffer, self.episode_len_buffer, self.episode_timeout_buffer)
               + self._stats(self.episode_reward_buffer_v, self.episode_len_buffer_v, self.episode_timeout_buffer_v)
               + [self.true_mean_reward, self.true_mean_reward_v, lr] + [summary.get(k, np.nan) for k in LOSS_KEYS])
        self.rows.append(row)
        if self.logdir:
            with open(os.path.join(self.logdir, "log-append.csv"), "a", newline="") as f:
                csv.writer(f).writerow(row)
        return dict(zip(self.columns, row))
