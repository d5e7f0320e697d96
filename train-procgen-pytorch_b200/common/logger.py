"""Episode statistics + CSV logging with the reference Logger's interface, accounting and column schema
(common/logger.py:13-200): ``feed(rew_batch, done_batch, true_mean_reward, rew_batch_v, done_batch_v,
true_mean_reward_v)`` once per iteration, ``dump(summary, lr)``, ``episode_reward_buffer``, ``logdir``, ``max_steps``.

Same accounting as the reference's double loop (:119-147): envs are walked in index order and, within an env, steps in
time order (ENV-MAJOR), every done flag closes the env's open episode (return = sum of its rewards since the previous
done, possibly begun in an earlier rollout; length likewise; time-out iff ``length == max_steps``), and the last 40
closed episodes live in the deques the statistics are taken from.  ``log-append.csv`` has the reference's columns in
the reference's order (time, 10 episode metrics, 10 ``val_`` metrics, ``ema_rewards``, the 9 loss metrics,
``learning_rate``).

Two producers feed it:
* ``feed(...)`` -- the reference call with [T, N] host batches; the O(T*N) Python loop is vectorised numpy.
* ``feed_episodes(...)`` -- the device path: ``Storage.snapshot_episodes`` runs ``tpp_episode_scan`` on the
  rollout in HBM and hands over only the episode count and the last 40 (return, length) records.
wandb / pandas are not used.
"""
from __future__ import annotations

import csv
import os
import time
import warnings
from collections import deque

import numpy as np

TIME_METRICS = ["timesteps", "wall_time", "num_episodes"]
LOSS_METRICS = ["loss_pi", "loss_v", "loss_entropy", "loss_x_entropy", "atn_entropy", "atn_entropy2",
                "loss_sparsity", "loss_feature_sparsity", "loss_total"]
EPISODE_METRICS = ["max_episode_rewards", "mean_episode_rewards", "median_episode_rewards", "min_episode_rewards",
                   "max_episode_len", "mean_episode_len", "min_episode_len", "mean_timeouts",
                   "mean_episode_len_pos_reward", "balanced_mean_rewards"]
KEEP = 40          # maxlen of the reference's episode deques (common/logger.py:33-35)


def close_episodes(rew, done, run_ret, run_len):
    """Vectorised restatement of the reference's episode walk.  rew, done: [T, N]; run_ret (float64 [N]) / run_len
    (int64 [N]) hold every env's open episode and are updated in place.  Returns (returns, lengths) of the episodes
    closed by this batch in env-major order."""
    rew = np.asarray(rew, dtype=np.float64)
    T, N = rew.shape
    csum = np.cumsum(rew, axis=0)
    e_idx, t_idx = np.nonzero(np.asarray(done).T > 0)          # row-major over [N, T] = env-major episode order
    prev_same = np.zeros(len(e_idx), dtype=bool)
    prev_same[1:] = e_idx[1:] == e_idx[:-1]
    start = np.where(prev_same, np.concatenate(([0], t_idx[:-1])), -1)       # previous done of the same env, or -1
    seg = csum[t_idx, e_idx] - np.where(start >= 0, csum[np.maximum(start, 0), e_idx], 0.0)
    rets = seg + np.where(start < 0, run_ret[e_idx], 0.0)
    lens = (t_idx - start) + np.where(start < 0, run_len[e_idx], 0)
    last = np.full(N, -1)
    last[e_idx] = t_idx                                          # (later entries of an env overwrite earlier ones)
    open_env = last < 0
    run_ret[:] = np.where(open_env, run_ret + csum[-1], csum[-1] - csum[np.maximum(last, 0), np.arange(N)])
    run_len[:] = np.where(open_env, run_len + T, T - 1 - last)
    return rets, lens


class Logger:
    def __init__(self, n_envs, logdir=None, use_wandb=False, has_vq=False, algo="ppo", greedy=False):
        self.start_time = time.time()
        self.n_envs, self.logdir, self.use_wandb, self.greedy = n_envs, logdir, use_wandb, greedy
        self.true_mean_reward = self.true_mean_reward_v = None
        self.episode_rewards = np.zeros(n_envs)                       # open-episode return / length per env
        self.episode_lens = np.zeros(n_envs, dtype=np.int64)
        self.episode_rewards_v = np.zeros(n_envs)
        self.episode_lens_v = np.zeros(n_envs, dtype=np.int64)
        self.episode_timeout_buffer = deque(maxlen=KEEP)
        self.episode_len_buffer = deque(maxlen=KEEP)
        self.episode_reward_buffer = deque(maxlen=KEEP)
        self.episode_timeout_buffer_v = deque(maxlen=KEEP)
        self.episode_len_buffer_v = deque(maxlen=KEEP)
        self.episode_reward_buffer_v = deque(maxlen=KEEP)
        self.max_steps = None            # set by the caller like the reference does (train.py:201)
        self.columns = (TIME_METRICS + EPISODE_METRICS + ["val_" + m for m in EPISODE_METRICS] + ["ema_rewards"]
                        + LOSS_METRICS + ["learning_rate"])
        self.timesteps, self.num_episodes = 0, 0
        self.log = []                    # rows of log-append.csv
        if logdir:
            os.makedirs(logdir, exist_ok=True)

    # ---- producers -------------------------------------------------------------------------------------------
    def _append(self, bufs, rets, lens):
        rew_buf, len_buf, to_buf = bufs
        for r, l in zip(rets[-KEEP:], lens[-KEEP:]):
            to_buf.append(1 if l == self.max_steps else 0)
            len_buf.append(int(l))
            rew_buf.append(r)

    def feed(self, rew_batch, done_batch, true_mean_reward=None, rew_batch_v=None, done_batch_v=None,
             true_mean_reward_v=None, *unused_greedy):
        self.true_mean_reward, self.true_mean_reward_v = true_mean_reward, true_mean_reward_v
        steps = rew_batch.shape[0]
        rets, lens = close_episodes(rew_batch, done_batch, self.episode_rewards, self.episode_lens)
        self._append((self.episode_reward_buffer, self.episode_len_buffer, self.episode_timeout_buffer), rets, lens)
        self.num_episodes += len(rets)
        if rew_batch_v is not None and done_batch_v is not None:
            rets, lens = close_episodes(rew_batch_v, done_batch_v, self.episode_rewards_v, self.episode_lens_v)
            self._append((self.episode_reward_buffer_v, self.episode_len_buffer_v, self.episode_timeout_buffer_v),
                         rets, lens)
        self.timesteps += self.n_envs * steps

    def feed_episodes(self, steps, record, true_mean_reward=None, record_v=None, true_mean_reward_v=None):
        """Device path: ``record`` = the float64 array ``tpp_episode_scan`` wrote (count, kept, (return, length) x kept);
        the open-episode state of every env stays on the device."""
        self.true_mean_reward, self.true_mean_reward_v = true_mean_reward, true_mean_reward_v
        for rec, bufs, train in ((record, (self.episode_reward_buffer, self.episode_len_buffer,
                                           self.episode_timeout_buffer), True),
                                 (record_v, (self.episode_reward_buffer_v, self.episode_len_buffer_v,
                                             self.episode_timeout_buffer_v), False)):
            if rec is None:
                continue
            total, kept = int(rec[0]), int(rec[1])
            pairs = np.asarray(rec[2:2 + 2 * kept]).reshape(kept, 2)
            self._append(bufs, pairs[:, 0], pairs[:, 1].astype(np.int64))
            if train:
                self.num_episodes += total
        self.timesteps += self.n_envs * steps

    # ---- statistics + CSV --------------------------------------------------------------------------------------
    @staticmethod
    def _episode_statistics(rew_buf, len_buf, to_buf, balanced):
        with warnings.catch_warnings():          # empty buffers give nan exactly like the reference's np.mean / np.median
            warnings.simplefilter("ignore")
            rew, ln = np.array(rew_buf, dtype=np.float64), np.array(len_buf)
            return [np.max(rew, initial=0), np.mean(rew), np.median(rew), np.min(rew, initial=0),
                    np.max(ln, initial=0), np.mean(ln), np.min(ln, initial=0), np.mean(np.array(to_buf)),
                    np.mean(ln[rew > 0]), balanced]

    def _get_episode_statistics(self):
        tr = self._episode_statistics(self.episode_reward_buffer, self.episode_len_buffer, self.episode_timeout_buffer,
                                      self.true_mean_reward)
        va = self._episode_statistics(self.episode_reward_buffer_v, self.episode_len_buffer_v,
                                      self.episode_timeout_buffer_v, self.true_mean_reward_v)
        return tr + va

    def dump(self, summary=None, lr=0.):
        summary = summary or {}
        wall_time = time.time() - self.start_time
        stats = self._get_episode_statistics()
        ema_reward = stats[1]
        if len(self.log) > 0:                                           # common/logger.py:154-157
            smoothing = .99 / (1 + len(self.log))
            prev_ema = self.log[-1][len(TIME_METRICS) + 2 * len(EPISODE_METRICS)]
            ema_reward = ema_reward * smoothing + prev_ema * (1 - smoothing)
        row = [self.timesteps, wall_time, self.num_episodes] + stats + [ema_reward] + list(summary.values()) + [lr]
        self.log.append(row)
        if self.logdir:
            with open(os.path.join(self.logdir, "log-append.csv"), "a", newline="") as f:
                writer = csv.writer(f)
                if f.tell() == 0:
                    writer.writerow(self.columns)
                writer.writerow(row)
        return dict(zip(self.columns, row))
