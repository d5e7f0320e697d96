"""ORACLE (test infrastructure, never imported by the product package).

Restatement of the reference's vectorised Box-World:

* grid transition / rewards / done       -> boxworld/box_world_env_vec.py:70-209
* level replacement + seed counter       -> boxworld/box_world_env_vec.py:204-207, 233-240, 272-297
* level generator                        -> boxworld/boxworld_gen_vec.py:4-25 (sampling_pairs), 40-97 (world_gen)
* reward/obs shaping wrappers            -> common/env/procgen_wrappers.py:282-355 (VecNormalize/RunningMeanStd),
                                            :391-419 (TransposeFrame, ScaledFloatFrame)

All integer state; the parity bar is bit-exact.  The generator draws from CPython's `random.Random`
(MT19937; `sample`/`choices`/`_randbelow` semantics are CPython 3.12's, which for the k=1 draws from a
set equal Python<=3.10's `sample(set, 1)` on `tuple(set)`; small-int sets iterate in ascending order, see
DESIGN.md).  Pinned by: the reference's ten scripted scenarios (boxworld/box_world_env_vec_test.py),
replayed in tests/test_oracle_golden.py, the committed fixtures tests/golden/boxworld_*.npz minted from the
live reference, and (container only) tests/test_oracle_vs_reference.py.
"""
from __future__ import annotations

import random

import numpy as np

COLORS = {1: [230, 190, 255], 2: [170, 255, 195], 3: [255, 250, 200], 4: [255, 216, 177], 5: [250, 190, 190],
          6: [240, 50, 230], 7: [145, 30, 180], 8: [67, 99, 216], 9: [66, 212, 244], 10: [60, 180, 75],
          11: [191, 239, 69], 12: [255, 255, 25], 13: [245, 130, 49], 14: [230, 25, 75], 15: [128, 0, 0],
          16: [154, 99, 36], 17: [128, 128, 0], 18: [70, 153, 144], 0: [0, 0, 117]}   # boxworld_gen_vec.py:28-32
NUM_COLORS = len(COLORS)
AGENT, GOAL, GRID, WALL = 128, 255, 220, 0          # grey levels of the four special colours (:34-38)
MOVES = np.array([[-1, 0], [1, 0], [0, -1], [0, 1]])  # UP, DOWN, LEFT, RIGHT (box_world_env_vec.py:60)


def sampling_pairs(rnd, num_pair, n):
    """boxworld_gen_vec.py:4-25 with the set population made explicit as an ascending tuple."""
    w = n - 1
    poss = set(range(1, n * w))
    keys, locks = [], []
    for _ in range(num_pair):
        key = rnd.sample(tuple(sorted(poss)), 1)[0]
        kx, ky = key // w, key % w
        gone = [kx * w + ky]
        gone += [kx * w + i + ky for i in range(1, min(2, n - 2 - ky) + 1)]
        gone += [kx * w - i + ky for i in range(1, min(2, ky) + 1)]
        poss -= set(gone)
        keys.append((kx, ky))
        locks.append((kx, ky + 1))
    agent = rnd.sample(tuple(sorted(poss)), 1)[0]
    poss.discard(agent)
    first = rnd.sample(tuple(sorted(poss)), 1)[0]
    return keys, locks, (first // w, first % w), (agent // w, agent % w)


def world_gen(n, goal_length, num_distractor, distractor_length, seed):
    """-> (world uint8 [n+2,n+2,3], player_position int64 [2], world_dic float64 [n+2,n+2]).
    boxworld_gen_vec.py:40-97."""
    rnd = random.Random(seed)
    dic = -np.ones((n + 2, n + 2))
    world = np.full((n + 2, n + 2, 3), GRID, dtype=np.uint8)
    world[0], world[-1], world[:, 0], world[:, -1] = WALL, WALL, WALL, WALL
    inner = world[1:-1, 1:-1]

    goal_cols = rnd.sample(range(NUM_COLORS), goal_length - 1)
    free_cols = [c for c in range(NUM_COLORS) if c not in goal_cols]
    dis_cols = [rnd.sample(free_cols, distractor_length) for _ in range(num_distractor)]
    dis_roots = rnd.choices(range(goal_length - 1), k=num_distractor)
    keys, locks, first_key, agent = sampling_pairs(rnd, goal_length - 1 + distractor_length * num_distractor, n)

    for i in range(1, goal_length):                                   # the goal path (:56-66)
        inner[keys[i - 1]] = [GOAL] * 3 if i == goal_length - 1 else COLORS[goal_cols[i]]
        inner[locks[i - 1]] = COLORS[goal_cols[i - 1]]
        dic[locks[i - 1][0] + 1, locks[i - 1][1] + 1] = 1
    inner[first_key] = COLORS[goal_cols[0]]                           # the loose first key (:69)
    for i, (dc, root) in enumerate(zip(dis_cols, dis_roots)):         # distractor branches (:74-87)
        kd = keys[goal_length - 1 + i * distractor_length: goal_length - 1 + (i + 1) * distractor_length]
        inner[kd[0][0], kd[0][1] + 1] = COLORS[goal_cols[root]]
        inner[kd[0]] = COLORS[dc[0]]
        dic[kd[0][0] + 1, kd[0][1] + 2] = 0
        for k, key in enumerate(kd[1:]):
            inner[key] = COLORS[dc[k - 1]]          # k == 0 wraps to the LAST colour (reference quirk)
            inner[key[0], key[1] + 1] = COLORS[dc[k]]
            dic[key[0] + 1, key[1] + 2] = 0
    inner[agent] = AGENT
    return world, np.array([agent[0] + 1, agent[1] + 1], dtype=np.int64), dic


class BoxWorldOracle:
    """box_world_env_vec.py:24-329, numpy, integer-exact."""

    def __init__(self, n_envs, n, goal_length, num_distractor, distractor_length, max_steps=10 ** 6,
                 start_seed=0, n_levels=0):
        self.num_envs, self.n = n_envs, n
        self.gen_args = (n, goal_length, num_distractor, distractor_length)
        self.max_steps, self.start_seed, self.n_levels = max_steps, start_seed, n_levels
        self.seed_counter = start_seed
        self.world = np.zeros((n_envs, n + 2, n + 2, 3), dtype=np.uint8)
        self.world_dic = np.zeros((n_envs, n + 2, n + 2))
        self.player_position = np.zeros((n_envs, 2), dtype=np.int64)
        self.owned_key = np.zeros((n_envs, 3), dtype=np.int64)
        self.num_env_steps = np.zeros(n_envs, dtype=np.int64)
        self.episode_reward = np.zeros(n_envs, dtype=np.int64)
        self.reward = np.zeros(n_envs, dtype=np.int64)
        self.done = np.zeros(n_envs, dtype=bool)
        self.solved = np.zeros(n_envs, dtype=bool)
        for i in range(n_envs):                      # start_envs (:272-292)
            self.replace_world_i(i, self.seed_counter)
            self._increment_seed()

    def _increment_seed(self):                       # :294-297
        self.seed_counter += 1
        if self.n_levels > 0:
            self.seed_counter = ((self.seed_counter - self.start_seed) % self.n_levels) + self.start_seed

    def replace_world_i(self, i, seed):              # :233-240
        self.world[i], self.player_position[i], self.world_dic[i] = world_gen(*self.gen_args, seed)
        self.num_env_steps[i] = 0
        self.episode_reward[i] = 0
        self.owned_key[i] = GRID

    def _cell(self, pos):
        p = np.clip(pos, 0, self.n + 1)
        return np.arange(self.num_envs), p[:, 0], p[:, 1]

    def step(self, action):
        action = np.asarray(action)
        w, e = self.world, np.arange(self.num_envs)
        cur = self.player_position.copy()
        new = cur + MOVES[action]
        self.num_env_steps += 1
        reward = np.zeros(self.num_envs, dtype=np.int64)
        done = self.num_env_steps == self.max_steps
        at, left, right = self._cell(new), self._cell(new - [0, 1]), self._cell(new + [0, 1])
        in_grid = np.all(new > 0, 1) & np.all(new <= self.n, 1)
        here, left_c, right_c = w[at].astype(np.int64), w[left].astype(np.int64), w[right].astype(np.int64)
        empty = np.all(here == GRID, 1)
        left_clear = (new[:, 1] == 1) | np.all(left_c == GRID, 1)
        first_key = ~empty & left_clear & (np.all(right_c == GRID, 1) | np.all(right_c == AGENT, 1))
        status = self.world_dic[at]
        is_lock = status != -1
        key_fits = np.all((self.owned_key == here) & (self.owned_key != GRID), 1)
        blocked = ~(empty | first_key | is_lock) | (is_lock & ~key_fits)
        settled = ~in_grid | blocked

        walk = empty & in_grid & ~settled                                    # :134-139
        take = first_key & in_grid & ~settled & ~walk                        # :141-149
        unlock = is_lock & key_fits & ~settled & ~walk & ~take               # :156-175
        for m in (walk, take, unlock):
            w[e[m], cur[m, 0], cur[m, 1]] = GRID
        w[e[unlock], left[1][unlock], left[2][unlock]] = GRID
        for m in (walk, take, unlock):
            w[e[m], at[1][m], at[2][m]] = AGENT
            self.player_position[m] = new[m]
        w[take, 0, 0] = here[take]
        self.owned_key[take] = here[take]
        reward[take] += 1
        w[unlock, 0, 0] = left_c[unlock]
        self.owned_key[unlock] = left_c[unlock]
        is_goal = unlock & np.all(left_c == GOAL, 1)
        reward[is_goal] += 10
        reward[unlock & (status == 1)] += 1
        is_distractor = unlock & (status == 0)
        reward[is_distractor] -= 1
        done = done | is_distractor | is_goal

        self.episode_reward += reward
        self.reward, self.done, self.solved = reward, done, is_goal
        self.moved_player = in_grid & ~blocked
        self.finished_return = self.episode_reward.copy()
        self.finished_length = self.num_env_steps.copy()
        for i in np.where(done)[0]:                                          # :204-207
            self.replace_world_i(i, self.seed_counter)
            self._increment_seed()
        return self.world, self.reward, self.done


# ----------------------------------------------------------------------------------------------
# Wrappers the reference stacks on BoxWorld / Procgen (create_box_world.py:51-67)
# ----------------------------------------------------------------------------------------------

class RunningMeanStdOracle:
    """procgen_wrappers.py:282-311 (scalar shape)."""

    def __init__(self, epsilon=1e-4):
        self.mean, self.var, self.count = 0.0, 1.0, epsilon

    def update(self, x):
        b_mean, b_var, b_n = np.mean(x), np.var(x), x.shape[0]
        delta = b_mean - self.mean
        tot = self.count + b_n
        new_mean = self.mean + delta * b_n / tot
        m2 = self.var * self.count + b_var * b_n + np.square(delta) * self.count * b_n / tot
        self.mean, self.var, self.count = new_mean, m2 / tot, tot


class VecNormalizeOracle:
    """Return-based reward scaling, procgen_wrappers.py:314-342 with ob=False."""

    def __init__(self, n_envs, gamma=0.99, cliprew=10.0, epsilon=1e-8):
        self.ret = np.zeros(n_envs)
        self.rms = RunningMeanStdOracle()
        self.gamma, self.cliprew, self.epsilon = gamma, cliprew, epsilon

    def step(self, rews, dones):
        self.ret = self.ret * self.gamma + rews
        self.rms.update(self.ret)
        out = np.clip(rews / np.sqrt(self.rms.var + self.epsilon), -self.cliprew, self.cliprew)
        self.ret[dones] = 0.0
        return out


def frame_to_obs(world):
    """TransposeFrame + ScaledFloatFrame (procgen_wrappers.py:391-419): NHWC uint8 -> NCHW float / 255."""
    return world.transpose(0, 3, 1, 2) / 255.0
