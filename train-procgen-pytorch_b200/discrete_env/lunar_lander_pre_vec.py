"""lunar_lander_pre_vec — vectorised lunar lander with this repository's OWN semantics (parity unpinned).

The reference names this family but has no implementation: `discrete_env/lunar_lander_pre_vec.py:16` raises
NotImplementedError at import, `common/env/env_constructor.py:27-28` raises for "lunar_lander", and the file body is
Gymnasium's scalar Box2D lander (a third-party physics engine that is not vendored).  What is kept from that file:
the 8-wide observation [x, y, vx, vy, angle, 20*omega/FPS, leg1, leg2] and its normalisation (:606-615), the four
actions {noop, left engine, main engine, right engine} with the same impulse geometry (:520-601), the shaping reward
(:617-633) and the +-100 terminal rewards (:635-642).  What is this repo's own model (oracle/lunar.py restates it):
one rigid hull with two massless legs, spring-damper foot contacts on flat ground at the helipad height, no engine
dispersion noise, dt = 1/50; "landed" = both feet down and |v|, |omega| < 0.05.
"""
import numpy as np

from .pre_vec_env import PreVecEnv, create_pre_vec

FPS, SCALE = 50.0, 30.0
_H2 = 400 / SCALE / 2
_START_Y = (400 / SCALE - ((400 / SCALE) / 4 + 18 / SCALE)) / _H2


class LunarLanderVecEnv(PreVecEnv):
    family = "lunar_lander"
    n_state = 8
    n_obs = 8
    envs_per_thread = 1     # ALU-bound (RK4 / contact model): occupancy beats vector width (profiles/README.md)

    def __init__(self, n_envs, gravity=10.0, main_engine_power=13.0, side_engine_power=0.6, initial_random=1000.0,
                 max_steps=1000, seed=0, drop_same=False, render_mode=None, device="cuda", numpy_compat=False):
        self.drop_same = drop_same
        self.gravity, self.main_engine_power, self.side_engine_power = gravity, main_engine_power, side_engine_power
        self.initial_random = initial_random
        dv = initial_random / FPS / 4.82                      # one step of the initial random force on the hull
        vx0, vy0 = dv * 10.0 / FPS, dv * _H2 / FPS
        self.high = np.array([1.5, 1.5, 5.0, 5.0, 3.14, 5.0, 1.0, 1.0], dtype=np.float32)   # Gymnasium's obs bounds
        self.low = -self.high
        self.low[6:] = 0.0
        self.start_low = [0.0, _START_Y, -vx0, -vy0, 0.0, 0.0, 0.0, 0.0]
        self.start_high = [0.0, _START_Y, vx0, vy0, 0.0, 0.0, 0.0, 0.0]
        self.kernel_params = [gravity, main_engine_power, side_engine_power]
        self.customizable_params = ["gravity", "main_engine_power", "side_engine_power", "initial_random", "max_steps"]
        super().__init__(n_envs, 4, "LunarLander", max_steps, seed, render_mode, device, numpy_compat)

    def get_ob_names(self):
        return ["x", "y", "vx", "vy", "angle", "angular velocity", "left leg contact", "right leg contact"]

    def get_action_lookup(self):
        return {0: "noop", 1: "fire left engine", 2: "fire main engine", 3: "fire right engine"}


def create_lunar_lander(args, hyperparameters, is_valid=False):
    param_range = {"gravity": [10.0, 11.0], "main_engine_power": [13.0], "side_engine_power": [0.6],
                   "initial_random": [1000.0, 1500.0]}
    return create_pre_vec(args, hyperparameters, param_range, LunarLanderVecEnv, is_valid)
