"""cartpole_swing: same dynamics as cartpole_pre_vec, pole starts hanging (theta0 ~ pi), done on |x| only,
shaped reward max(cos theta, 0) * cos(pi x / (2 x_threshold)) on the new state
(reference: discrete_env/cartpole_swing_pre_vec.py:19-239, 313-327)."""
import numpy as np

from .cartpole_pre_vec import CARTPOLE_PARAM_RANGE
from .pre_vec_env import PreVecEnv, create_pre_vec


class CartPoleSwingVecEnv(PreVecEnv):
    family = "cartpole_swing"
    n_state = 9
    n_obs = 9

    def __init__(self, n_envs, h_range=2.4, min_gravity=9.8, max_gravity=10.4, min_pole_length=0.5,
                 max_pole_length=1.0, min_cart_mass=1.0, max_cart_mass=1.5, min_pole_mass=0.1, max_pole_mass=0.2,
                 min_force_mag=10., max_force_mag=10., max_steps=1000, seed=0, drop_same=False, render_mode=None,
                 device="cuda", numpy_compat=False):
        self.drop_same = drop_same
        for k, v in dict(min_gravity=min_gravity, max_gravity=max_gravity, min_cart_mass=min_cart_mass,
                         max_cart_mass=max_cart_mass, min_pole_mass=min_pole_mass, max_pole_mass=max_pole_mass,
                         min_pole_length=min_pole_length, max_pole_length=max_pole_length,
                         min_force_mag=min_force_mag, max_force_mag=max_force_mag, h_range=h_range).items():
            setattr(self, k, v)
        self.tau = 0.02
        self.kinematics_integrator = "euler"
        self.x_threshold = h_range
        fmax = np.finfo(np.float32).max
        self.high = np.array([fmax] * 4 + [max_gravity, max_pole_length, max_cart_mass, max_pole_mass, max_force_mag],
                             dtype=np.float32)
        self.low = -self.high
        self.low[4:] = [min_gravity, min_pole_length, min_cart_mass, min_pole_mass, min_force_mag]
        self.start_low = [-0.05, -0.05, np.pi - 0.05, -0.05, min_gravity, min_pole_length, min_cart_mass,
                          min_pole_mass, min_force_mag]
        self.start_high = [0.05, 0.05, np.pi + 0.05, 0.05, max_gravity, max_pole_length, max_cart_mass, max_pole_mass,
                           max_force_mag]
        self.kernel_params = [self.x_threshold, 0.0, self.tau]
        self.customizable_params = ["h_range", "min_gravity", "max_gravity", "max_steps", "min_cart_mass",
                                    "max_cart_mass", "min_pole_mass", "max_pole_mass", "min_pole_length",
                                    "max_pole_length", "min_force_mag", "max_force_mag", "tau",
                                    "kinematics_integrator"]
        super().__init__(n_envs, 2, "CartPole", max_steps, seed, render_mode, device, numpy_compat)

    def get_ob_names(self):
        return ["Cart Position", "Cart Velocity", "Pole Angle", "Pole Angular Velocity", "Gravity", "Pole Length",
                "Cart Mass", "Pole Mass", "Action Force"]

    def get_action_lookup(self):
        return {0: "push left", 1: "push right"}


def create_cartpole_swing(args, hyperparameters, is_valid=False):
    rng = {k: v for k, v in CARTPOLE_PARAM_RANGE.items() if k != "degrees"}
    return create_pre_vec(args, hyperparameters, rng, CartPoleSwingVecEnv, is_valid)
