"""ORACLE (test infrastructure, never imported by the product package).

float64 numpy restatement of the reference's pre-vectorised env hot path:

* step / truncate / auto-reset contract      -> discrete_env/pre_vec_env.py:78-93, 98-119
* start-space sampling (PCG64, full block)    -> discrete_env/helper_pre_vec.py:28-38, pre_vec_env.py:121-126
* cartpole dynamics / done                    -> discrete_env/cartpole_pre_vec.py:210-256
* cartpole swing-up dynamics / reward         -> discrete_env/cartpole_swing_pre_vec.py:195-239
* mountain car dynamics + rejection start     -> discrete_env/mountain_car_pre_vec.py:158-161, 194-209
* acrobot RK4 "book" dynamics, wrap, bound    -> discrete_env/acrobot_pre_vec.py:279-394, 450-541

Pinned against the live reference (tests/test_oracle_vs_reference.py, container only) and against the
committed fixtures minted from it (tests/golden/prevec_*.npz, oracle/mint_golden.py).  When seeded the
same way it reproduces the reference's free-running trajectories bit for bit (same PCG64 stream, same
float64 operation order), except for the acrobot `wrap()` cross-env bug (acrobot_pre_vec.py:467-468),
where the oracle wraps each env independently and theta is compared mod 2*pi.

Every function is pure numpy; the stateful `OraclePreVec` wrapper only carries (state, n_steps, rng).
"""
from __future__ import annotations

import math

import numpy as np

# ----------------------------------------------------------------------------------------------
# Family descriptions: start-space bounds follow the ctor defaults of each reference class;
# `create_*` param ranges (train / `_v` validation) are in FAMILY_RANGES below.
# ----------------------------------------------------------------------------------------------

CARTPOLE_DEFAULTS = dict(
    degrees=12, h_range=2.4, min_gravity=9.8, max_gravity=10.4, min_pole_length=0.5, max_pole_length=1.0,
    min_cart_mass=1.0, max_cart_mass=1.5, min_pole_mass=0.1, max_pole_mass=0.2, min_force_mag=10.0,
    max_force_mag=10.0, max_steps=500)
SWING_DEFAULTS = dict(CARTPOLE_DEFAULTS, max_steps=1000)
SWING_DEFAULTS.pop("degrees")
MOUNTAIN_CAR_DEFAULTS = dict(
    goal_velocity=0.0, left_boundary=-1.2, min_start_position=-0.6, max_start_position=-0.4, max_speed=0.07,
    min_goal_position=0.5, max_goal_position=3.0, min_gravity=0.001, max_gravity=0.0025,
    min_right_boundary=0.6, max_right_boundary=5.0, force=0.001, max_steps=500, sparse_rewards=True)
ACROBOT_DEFAULTS = dict(
    gravity=(9.8, 11.4), link_length_1=(1.0, 1.5), link_length_2=(1.0, 1.5), link_mass_1=(1.0, 1.5),
    link_mass_2=(1.0, 1.5), link_com_pos_1=(0.5, 0.5), link_com_pos_2=(0.5, 0.5), link_moi=(1.0, 1.0),
    max_vel_1=4 * math.pi, max_vel_2=9 * math.pi, max_steps=500)

TAU = 0.02          # cartpole_pre_vec.py:127
ACROBOT_DT = 0.2    # acrobot_pre_vec.py:147


def cartpole_accel(state, action):
    """Shared cart-pole accelerations (cartpole_pre_vec.py:210-229). Returns (xacc, thetaacc)."""
    x, x_dot, theta, theta_dot, g, length, m_c, m_p, f_mag = state.T
    force = np.where(np.asarray(action) == 0, -1.0, 1.0) * f_mag
    c, s = np.cos(theta), np.sin(theta)
    pml = m_p * length
    total = m_p + m_c
    temp = (force + pml * theta_dot ** 2 * s) / total
    thetaacc = (g * s - c * temp) / (length * (4.0 / 3.0 - m_p * c ** 2 / total))
    xacc = temp - pml * thetaacc * c / total
    return xacc, thetaacc


def _cartpole_euler(state, action):
    xacc, thetaacc = cartpole_accel(state, action)
    ns = state.copy()
    ns[:, 0] = state[:, 0] + TAU * state[:, 1]
    ns[:, 1] = state[:, 1] + TAU * xacc
    ns[:, 2] = state[:, 2] + TAU * state[:, 3]
    ns[:, 3] = state[:, 3] + TAU * thetaacc
    return ns


def cartpole_transition(state, action, x_threshold=2.4, theta_threshold=12 * 2 * math.pi / 360):
    """-> (next_state, terminated, reward). cartpole_pre_vec.py:210-256 (reward == 1 always, :114)."""
    ns = _cartpole_euler(state, action)
    x, th = ns[:, 0], ns[:, 2]
    term = (x < -x_threshold) | (x > x_threshold) | (th < -theta_threshold) | (th > theta_threshold)
    return ns, term, np.ones(len(ns))


def swing_transition(state, action, x_threshold=2.4):
    """cartpole_swing_pre_vec.py:195-239: done on |x| only, shaped reward on the NEW state."""
    ns = _cartpole_euler(state, action)
    x, th = ns[:, 0], ns[:, 2]
    term = (x < -x_threshold) | (x > x_threshold)
    r_theta = np.cos(th)
    r_theta[r_theta < 0] = 0
    r_x = np.cos((x / x_threshold) * (np.pi / 2.0))
    return ns, term, r_theta * r_x


def mountain_car_transition(state, action, force=0.001, max_speed=0.07, left_boundary=-1.2, goal_velocity=0.0,
                            sparse_rewards=True):
    """mountain_car_pre_vec.py:194-209."""
    pos, vel, g, right, goal = (c.copy() for c in state.T)
    vel = vel + ((np.asarray(action) - 1) * force + np.cos(3 * pos) * (-g))
    vel = np.clip(vel, -max_speed, max_speed)
    pos = pos + vel
    pos = np.clip(pos, left_boundary, right)
    vel[(pos == left_boundary) & (vel < 0)] = 0
    term = (pos >= goal) & (vel >= goal_velocity)
    ns = np.stack((pos, vel, g, right, goal), axis=1)
    if sparse_rewards:
        rew = np.full(len(ns), -1.0)
    else:  # _height(position) - 1, mountain_car_pre_vec.py:222-223
        rew = np.sin(3 * pos) * 0.45 + 0.55 - 1
    return ns, term, rew


def acrobot_dsdt(s_aug, params):
    """acrobot_pre_vec.py:359-394 ("book" branch). s_aug = [th1, th2, dth1, dth2, torque]."""
    g, l1, l2, m1, m2, lc1, lc2, moi = params.T
    i1 = i2 = moi
    a = s_aug[:, 4]
    th1, th2, dth1, dth2 = s_aug[:, 0], s_aug[:, 1], s_aug[:, 2], s_aug[:, 3]
    d1 = m1 * lc1 ** 2 + m2 * (l1 ** 2 + lc2 ** 2 + 2 * l1 * lc2 * np.cos(th2)) + i1 + i2
    d2 = m2 * (lc2 ** 2 + l1 * lc2 * np.cos(th2)) + i2
    phi2 = m2 * lc2 * g * np.cos(th1 + th2 - np.pi / 2.0)
    phi1 = (-m2 * l1 * lc2 * dth2 ** 2 * np.sin(th2)
            - 2 * m2 * l1 * lc2 * dth2 * dth1 * np.sin(th2)
            + (m1 * lc1 + m2 * l1) * g * np.cos(th1 - np.pi / 2)
            + phi2)
    ddth2 = (a + d2 / d1 * phi1 - m2 * l1 * lc2 * dth1 ** 2 * np.sin(th2) - phi2) / (m2 * lc2 ** 2 + i2 - d2 ** 2 / d1)
    ddth1 = -(d2 * ddth2 + phi1) / d1
    return np.stack((dth1, dth2, ddth1, ddth2, np.zeros_like(dth1)), axis=1)


def wrap_per_env(x, m, M):
    """Per-env-correct version of acrobot_pre_vec.py:450-469 (the reference couples envs)."""
    x = x.copy()
    diff = M - m
    while np.any(x > M):
        x[x > M] -= diff
    while np.any(x < m):
        x[x < m] += diff
    return x


def acrobot_transition(state, action, max_vel_1=4 * math.pi, max_vel_2=9 * math.pi):
    """acrobot_pre_vec.py:279-307: one RK4 step (dt=0.2), wrap, bound, terminal, reward."""
    torque = np.array([-1.0, 0.0, 1.0])[np.asarray(action)]
    params = state[:, 4:]
    y0 = np.concatenate((state[:, :4], torque[:, None]), axis=1)
    dt, dt2 = ACROBOT_DT, ACROBOT_DT / 2.0
    k1 = acrobot_dsdt(y0, params)
    k2 = acrobot_dsdt(y0 + dt2 * k1, params)
    k3 = acrobot_dsdt(y0 + dt2 * k2, params)
    k4 = acrobot_dsdt(y0 + dt * k3, params)
    ns4 = (y0 + dt / 6.0 * (k1 + 2 * k2 + 2 * k3 + k4))[:, :4]
    ns = state.copy()
    ns[:, 0] = wrap_per_env(ns4[:, 0], -np.pi, np.pi)
    ns[:, 1] = wrap_per_env(ns4[:, 1], -np.pi, np.pi)
    ns[:, 2] = np.clip(ns4[:, 2], -max_vel_1, max_vel_1)
    ns[:, 3] = np.clip(ns4[:, 3], -max_vel_2, max_vel_2)
    term = -np.cos(ns[:, 0]) - np.cos(ns[:, 1] + ns[:, 0]) > 1.0
    rew = np.where(term, 0.0, -1.0)
    return ns, term, rew


def acrobot_obs(state):
    """acrobot_pre_vec.py:309-319: [cos th1, sin th1, cos th2, sin th2, dth1, dth2, 8 params]."""
    return np.concatenate((np.cos(state[:, 0:1]), np.sin(state[:, 0:1]), np.cos(state[:, 1:2]),
                           np.sin(state[:, 1:2]), state[:, 2:]), axis=1)


# ----------------------------------------------------------------------------------------------
# Start spaces
# ----------------------------------------------------------------------------------------------

def start_bounds(family, **kw):
    """(low, high) of the reference StartSpace for a family, kwargs named like the reference ctor."""
    if family in ("cartpole", "cartpole_swing"):
        p = dict(CARTPOLE_DEFAULTS if family == "cartpole" else SWING_DEFAULTS, **kw)
        c = 0.0 if family == "cartpole" else np.pi
        low = [-0.05, -0.05, c - 0.05, -0.05, p["min_gravity"], p["min_pole_length"], p["min_cart_mass"],
               p["min_pole_mass"], p["min_force_mag"]]
        high = [0.05, 0.05, c + 0.05, 0.05, p["max_gravity"], p["max_pole_length"], p["max_cart_mass"],
                p["max_pole_mass"], p["max_force_mag"]]
    elif family == "mountain_car":
        p = dict(MOUNTAIN_CAR_DEFAULTS, **kw)
        low = [p["min_start_position"], 0, p["min_gravity"], p["min_right_boundary"], p["min_goal_position"]]
        high = [p["max_start_position"], 0, p["max_gravity"], p["max_right_boundary"], p["max_goal_position"]]
    elif family == "acrobot":
        p = dict(ACROBOT_DEFAULTS, **kw)
        ctx = [p[k] for k in ("gravity", "link_length_1", "link_length_2", "link_mass_1", "link_mass_2",
                              "link_com_pos_1", "link_com_pos_2", "link_moi")]
        low = [-0.1] * 4 + [c[0] for c in ctx]
        high = [0.1] * 4 + [c[-1] for c in ctx]
    else:
        raise KeyError(family)
    return np.array(low, dtype=np.float64), np.array(high, dtype=np.float64)


def sample_start(rng, low, high, n, reject=None):
    """helper_pre_vec.py:28-38: one full (n, n_state) uniform block; with a rejection condition draw
    10*n rows at a time and keep the first n accepted."""
    if reject is None:
        return rng.uniform(low=low, high=high, size=(n, len(low)))
    x = rng.uniform(low=low, high=high, size=(n * 10, len(low)))
    x = x[~reject(x)]
    while len(x) < n:
        z = rng.uniform(low=low, high=high, size=(n * 10, len(low)))
        x = np.vstack((x, z[~reject(z)]))
    return x[:n]


def mountain_car_reject(x):
    return x[:, 4] > x[:, 3]  # goal_position > right_boundary (mountain_car_pre_vec.py:161)


# ----------------------------------------------------------------------------------------------
# The step contract
# ----------------------------------------------------------------------------------------------

class OraclePreVec:
    """pre_vec_env.py:32-126 as a small state machine.

    `step(action, reset_rows=None)`: when `reset_rows` ([N, n_state]) is given those rows are used for the
    finished envs (teacher forcing for the CUDA parity tests); otherwise a full block is drawn from the
    PCG64 stream exactly as the reference does (`np.any(terminated)` guard included).
    """

    def __init__(self, family, n_envs, seed=0, **kw):
        if n_envs < 2:
            raise Exception("n_envs must be greater than or equal to 2")
        self.family, self.n_envs = family, n_envs
        defaults = {"cartpole": CARTPOLE_DEFAULTS, "cartpole_swing": SWING_DEFAULTS,
                    "mountain_car": MOUNTAIN_CAR_DEFAULTS, "acrobot": ACROBOT_DEFAULTS}[family]
        self.p = dict(defaults, **kw)
        self.max_steps = self.p["max_steps"]
        self.low, self.high = start_bounds(family, **kw)
        self.reject = mountain_car_reject if family == "mountain_car" else None
        self.state = np.zeros((n_envs, len(self.low)))
        self.n_steps = np.zeros(n_envs)
        self.terminated = np.full(n_envs, True)
        self.reward = np.zeros(n_envs)
        self.rng = None
        self.reset(seed=seed)

    # -- reference: PreVecEnv.seed / set / reset ------------------------------------------------
    def seed(self, seed=None):
        if seed is not None:
            self.rng = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))

    def sample_block(self):
        return sample_start(self.rng, self.low, self.high, self.n_envs, self.reject)

    def _set(self, rows):
        self.state[self.terminated] = rows[self.terminated]
        self.n_steps[self.terminated] = 0

    def reset(self, seed=None, rows=None):
        self.terminated = np.full(self.n_envs, True)
        self.n_steps = np.zeros(self.n_envs)
        self.seed(seed)
        self._set(self.sample_block() if rows is None else rows)
        return self.obs()

    # -- reference: PreVecEnv.step ----------------------------------------------------------------
    def transition(self, action):
        p = self.p
        if self.family == "cartpole":
            return cartpole_transition(self.state, action, p["h_range"], p["degrees"] * 2 * math.pi / 360)
        if self.family == "cartpole_swing":
            return swing_transition(self.state, action, p["h_range"])
        if self.family == "mountain_car":
            return mountain_car_transition(self.state, action, p["force"], p["max_speed"], p["left_boundary"],
                                           p["goal_velocity"], p["sparse_rewards"])
        return acrobot_transition(self.state, action, p["max_vel_1"], p["max_vel_2"])

    def step(self, action, reset_rows=None):
        action = np.asarray(action)
        assert action.size == self.n_envs
        self.state, self.terminated, self.reward = self.transition(action)
        self.n_steps += 1
        self.terminated = self.terminated | (self.n_steps >= self.max_steps)
        self.pre_reset_state = self.state.copy()
        if np.any(self.terminated):
            self._set(self.sample_block() if reset_rows is None else reset_rows)
        return self.obs(), self.reward, self.terminated.copy()

    def obs(self):
        return acrobot_obs(self.state) if self.family == "acrobot" else self.state.copy()
