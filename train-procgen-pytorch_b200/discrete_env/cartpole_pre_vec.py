"""cartpole_pre_vec: 9-column cart-pole with per-env randomised physics (reference:
discrete_env/cartpole_pre_vec.py:20-256, 397-412).  State == observation:
[x, x_dot, theta, theta_dot, gravity, pole_length, cart_mass, pole_mass, force_mag]."""
import math

import numpy as np

from .pre_vec_env import PreVecEnv, create_pre_vec


class CartPoleVecEnv(PreVecEnv):
    family = "cartpole"
    n_state = 9
    n_obs = 9

    def __init__(self, n_envs, degrees=12, h_range=2.4, min_gravity=9.8, max_gravity=10.4, min_pole_length=0.5,
                 max_pole_length=1.0, min_cart_mass=1.0, max_cart_mass=1.5, min_pole_mass=0.1, max_pole_mass=0.2,
                 min_force_mag=10., max_force_mag=10., max_steps=500, seed=0, continuous=False, drop_same=False,
                 render_mode=None, device="cuda", numpy_compat=False):
        if continuous:
            raise NotImplementedError("continuous-action cartpole is outside the north-star hot path")
        self.continuous = continuous
        self.drop_same = drop_same
        for k, v in dict(min_gravity=min_gravity, max_gravity=max_gravity, min_cart_mass=min_cart_mass,
                         max_cart_mass=max_cart_mass, min_pole_mass=min_pole_mass, max_pole_mass=max_pole_mass,
                         min_pole_length=min_pole_length, max_pole_length=max_pole_length,
                         min_force_mag=min_force_mag, max_force_mag=max_force_mag, degrees=degrees,
                         h_range=h_range).items():
            setattr(self, k, v)
        self.tau = 0.02
        self.kinematics_integrator = "euler"
        self.theta_threshold_radians = degrees * 2 * math.pi / 360
        self.x_threshold = h_range
        fmax = np.finfo(np.float32).max
        self.high = np.array([self.x_threshold * 2, fmax, self.theta_threshold_radians * 2, fmax, max_gravity,
                              max_pole_length, max_cart_mass, max_pole_mass, max_force_mag], dtype=np.float32)
        self.low = -self.high
        self.low[4:] = [min_gravity, min_pole_length, min_cart_mass, min_pole_mass, min_force_mag]
        self.start_low = [-0.05] * 4 + [min_gravity, min_pole_length, min_cart_mass, min_pole_mass, min_force_mag]
        self.start_high = [0.05] * 4 + [max_gravity, max_pole_length, max_cart_mass, max_pole_mass, max_force_mag]
        self.kernel_params = [self.x_threshold, self.theta_threshold_radians, self.tau]
        self.customizable_params = ["degrees", "h_range", "min_gravity", "max_gravity", "max_steps", "min_cart_mass",
                                    "max_cart_mass", "min_pole_mass", "max_pole_mass", "min_pole_length",
                                    "max_pole_length", "min_force_mag", "max_force_mag", "tau",
                                    "kinematics_integrator"]
        super().__init__(n_envs, 2, "CartPole", max_steps, seed, render_mode, device, numpy_compat)

    def get_ob_names(self):
        return ["Cart Position", "Cart Velocity", "Pole Angle", "Pole Angular Velocity", "Gravity", "Pole Length",
                "Cart Mass", "Pole Mass", "Action Force"]

    def get_action_lookup(self):
        return {0: "push left", 1: "push right"}

    def rew_func(self, state):
        return np.ones((*state.shape[:-1], 1))

    def done_func(self, state):
        x, theta = state[..., 0], state[..., 2]
        return (x < -self.x_threshold) | (x > self.x_threshold) | (theta < -self.theta_threshold_radians) | \
               (theta > self.theta_threshold_radians)


CARTPOLE_PARAM_RANGE = {
    "degrees": [12], "h_range": [2.4], "min_gravity": [9.8, 10.4], "max_gravity": [10.4, 24.8],
    "min_pole_length": [0.5, 1.0], "max_pole_length": [1.0, 2.0], "min_cart_mass": [1.0, 2.],
    "max_cart_mass": [1.5, 3.], "min_pole_mass": [0.1, .2], "max_pole_mass": [0.2, .4], "min_force_mag": [10.],
    "max_force_mag": [10.],
}


def create_cartpole(args, hyperparameters, is_valid=False):
    return create_pre_vec(args, hyperparameters, CARTPOLE_PARAM_RANGE, CartPoleVecEnv, is_valid)
