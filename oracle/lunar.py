"""ORACLE for lunar_lander_pre_vec — **parity unpinned**.

The reference has NO implementation of this family: `discrete_env/lunar_lander_pre_vec.py:16` raises
NotImplementedError at import and the rest of the file is Gymnasium's scalar Box2D lander (Box2D is not vendored and
not installable here).  This file is therefore a float64 restatement of THIS repository's own vectorised semantics
(csrc/env_prevec.cu, family TPP_LUNAR_LANDER); it checks the CUDA kernel against an independent implementation of
the same equations, not against the reference.  What is taken from the reference file: the 8-wide observation
layout and its normalisation (:615-624), the four discrete actions and engine impulse geometry (:520-608), the
reward shaping formula (:628-637) and the terminal rewards (:644-651); the observation and shaping formulas are
executed from the reference's own source lines in tests/test_oracle_vs_reference.py.

Model: one rigid body (hull) with two massless legs; flat ground at the helipad height across
the whole width; feet are penalty (spring-damper) contacts; no engine dispersion noise; dt = 1/50.
"""
from __future__ import annotations

import numpy as np

FPS, SCALE = 50.0, 30.0
W2, H2 = 600 / SCALE / 2, 400 / SCALE / 2                 # half world width / height
HELIPAD_Y = (400 / SCALE) / 4
LEG_DOWN = 18 / SCALE
MAIN_POWER, SIDE_POWER = 13.0, 0.6
MASS, INERTIA, GRAVITY = 4.82, 0.84, 10.0
FOOT = np.array([[-20 / SCALE, -26 / SCALE], [20 / SCALE, -26 / SCALE]])
HULL = np.array([[-17 / SCALE, -10 / SCALE], [17 / SCALE, -10 / SCALE]])
K_N, C_N, C_T, MU = 1500.0, 60.0, 30.0, 1.0
DT = 1.0 / FPS
START_LOW = np.array([0.0, (400 / SCALE - (HELIPAD_Y + LEG_DOWN)) / H2, -0.83, -0.553, 0.0, 0.0, 0.0, 0.0])
START_HIGH = np.array([0.0, (400 / SCALE - (HELIPAD_Y + LEG_DOWN)) / H2, 0.83, 0.553, 0.0, 0.0, 0.0, 0.0])


def observation(x, y, vx, vy, ang, om, c0, c1):
    """The 8-wide normalised observation of the reference file (discrete_env/lunar_lander_pre_vec.py:615-624)."""
    return np.stack([(x - W2) / W2, (y - (HELIPAD_Y + LEG_DOWN)) / H2, vx * W2 / FPS, vy * H2 / FPS, ang,
                     20.0 * om / FPS, c0, c1], axis=1)


def shaping(s):
    return (-100 * np.sqrt(s[:, 0] ** 2 + s[:, 1] ** 2) - 100 * np.sqrt(s[:, 2] ** 2 + s[:, 3] ** 2)
            - 100 * np.abs(s[:, 4]) + 10 * s[:, 6] + 10 * s[:, 7])


def lunar_transition(state, action):
    """state [N, 8] normalised observation (== state), action [N] in {0 noop, 1 left, 2 main, 3 right}.
    -> (next_state, terminated, reward)."""
    s = np.asarray(state, dtype=np.float64)
    a = np.asarray(action)
    x, y = s[:, 0] * W2 + W2, s[:, 1] * H2 + (HELIPAD_Y + LEG_DOWN)
    vx, vy = s[:, 2] * FPS / W2, s[:, 3] * FPS / H2
    ang, om = s[:, 4].copy(), s[:, 5] * FPS / 20.0
    prev_shaping = shaping(s)
    sn, cs = np.sin(ang), np.cos(ang)
    tipx, tipy = sn, cs
    sidex, sidey = -tipy, tipx
    main = (a == 2).astype(np.float64)
    ox, oy = tipx * (4 / SCALE), -tipy * (4 / SCALE)
    jx, jy = -ox * MAIN_POWER * main, -oy * MAIN_POWER * main
    vx, vy = vx + jx / MASS, vy + jy / MASS
    om = om + (ox * jy - oy * jx) / INERTIA
    side = ((a == 1) | (a == 3)).astype(np.float64)
    d = (a - 2).astype(np.float64) * side
    ox, oy = sidex * (d * 12 / SCALE), -sidey * (d * 12 / SCALE)
    jx, jy = -ox * SIDE_POWER, -oy * SIDE_POWER
    rx, ry = ox - tipx * 17 / SCALE, oy + tipy * 14 / SCALE
    vx, vy = vx + jx / MASS, vy + jy / MASS
    om = om + (rx * jy - ry * jx) / INERTIA * side
    # foot contacts (penalty method) evaluated at the pre-integration pose
    fx_tot, fy_tot, tq = np.zeros_like(x), np.zeros_like(x), np.zeros_like(x)
    for f in range(2):
        rx = cs * FOOT[f, 0] - sn * FOOT[f, 1]
        ry = sn * FOOT[f, 0] + cs * FOOT[f, 1]
        pen = HELIPAD_Y - (y + ry)
        on = pen > 0
        vfx, vfy = vx - om * ry, vy + om * rx
        fn = np.where(on, np.maximum(K_N * pen - C_N * vfy, 0.0), 0.0)
        ft = np.where(on, np.clip(-C_T * vfx, -MU * fn, MU * fn), 0.0)
        fx_tot, fy_tot, tq = fx_tot + ft, fy_tot + fn, tq + rx * fn - ry * ft
    vx = vx + fx_tot * DT / MASS
    vy = vy + (fy_tot / MASS - GRAVITY) * DT
    om = om + tq * DT / INERTIA
    x, y, ang = x + vx * DT, y + vy * DT, ang + om * DT
    sn, cs = np.sin(ang), np.cos(ang)
    c = []
    for f in range(2):
        ry = sn * FOOT[f, 0] + cs * FOOT[f, 1]
        c.append((y + ry <= HELIPAD_Y).astype(np.float64))
    hull_low = np.minimum(y + sn * HULL[0, 0] + cs * HULL[0, 1], y + sn * HULL[1, 0] + cs * HULL[1, 1])
    ns = observation(x, y, vx, vy, ang, om, c[0], c[1])
    reward = shaping(ns) - prev_shaping - 0.30 * main - 0.03 * side
    crashed = (hull_low <= HELIPAD_Y) | (np.abs(ns[:, 0]) >= 1.0)
    landed = (c[0] > 0) & (c[1] > 0) & (np.abs(vx) < 0.05) & (np.abs(vy) < 0.05) & (np.abs(om) < 0.05) & ~crashed
    reward = np.where(crashed, -100.0, reward)
    reward = np.where(landed, 100.0, reward)
    return ns, crashed | landed, reward
