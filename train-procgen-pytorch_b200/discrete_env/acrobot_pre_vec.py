"""acrobot_pre_vec: 12-column state [theta1, theta2, dtheta1, dtheta2, gravity, l1, l2, m1, m2, lc1, lc2, moi],
one RK4 step (dt 0.2) of the "book" dynamics, 14-wide observation [cos/sin theta1, cos/sin theta2, dtheta1,
dtheta2, 8 params], reward -1 until terminal (reference: discrete_env/acrobot_pre_vec.py:31-394, 450-569).

The reference's `wrap()` couples envs (acrobot_pre_vec.py:467-468); this kernel wraps each env
independently, so raw angles agree with the reference modulo 2*pi and observations agree exactly."""
from math import pi

import numpy as np

from .pre_vec_env import PreVecEnv, create_pre_vec


class AcrobotVecEnv(PreVecEnv):
    family = "acrobot"
    n_state = 12
    n_obs = 14
    envs_per_thread = 1     # ALU-bound (RK4 / contact model): occupancy beats vector width (profiles/README.md)
    dt = 0.2
    AVAIL_TORQUE = np.array([-1.0, 0.0, +1])
    book_or_nips = "book"

    def __init__(self, n_envs, torque_noise_max=0.0, gravity=None, link_length_1=None, link_length_2=None,
                 link_mass_1=None, link_mass_2=None, link_com_pos_1=None, link_com_pos_2=None, link_moi=None,
                 max_vel_1=4 * pi, max_vel_2=9 * pi, max_steps=500, unprocessed_features=False, seed=0,
                 drop_same=False, render_mode=None, device="cuda", numpy_compat=False):
        if torque_noise_max > 0 or unprocessed_features:
            raise NotImplementedError("torque noise / unprocessed features are off in every reference config")
        self.drop_same = drop_same
        self.unprocessed_features = unprocessed_features
        self.torque_noise_max = torque_noise_max
        self.gravity = gravity if gravity is not None else [9.8, 11.4]
        self.link_length_1 = link_length_1 if link_length_1 is not None else [1.0, 1.5]
        self.link_length_2 = link_length_2 if link_length_2 is not None else [1.0, 1.5]
        self.link_mass_1 = link_mass_1 if link_mass_1 is not None else [1.0, 1.5]
        self.link_mass_2 = link_mass_2 if link_mass_2 is not None else [1.0, 1.5]
        self.link_com_pos_1 = link_com_pos_1 if link_com_pos_1 is not None else [0.5, 0.5]
        self.link_com_pos_2 = link_com_pos_2 if link_com_pos_2 is not None else [0.5, 0.5]
        self.link_moi = link_moi if link_moi is not None else [1.0, 1.0]
        self.max_vel_1, self.max_vel_2 = max_vel_1, max_vel_2
        ctx = np.array([self.gravity, self.link_length_1, self.link_length_2, self.link_mass_1, self.link_mass_2,
                        self.link_com_pos_1, self.link_com_pos_2, self.link_moi], dtype=np.float64)
        high = np.array([1.0, 1.0, 1.0, 1.0, max_vel_1, max_vel_2], dtype=np.float32)
        self.high = np.concatenate((high, ctx[:, -1])).astype(np.float32)
        self.low = np.concatenate((-high, ctx[:, 0])).astype(np.float32)
        self.start_low = [-0.1] * 4 + list(ctx[:, 0])
        self.start_high = [0.1] * 4 + list(ctx[:, -1])
        self.kernel_params = [max_vel_1, max_vel_2, self.dt]
        self.customizable_params = ["torque_noise_max", "gravity", "link_length_1", "link_length_2", "link_mass_1",
                                    "link_mass_2", "link_com_pos_1", "link_com_pos_2", "link_moi", "max_vel_1",
                                    "max_vel_2", "max_steps"]
        super().__init__(n_envs, 3, "Acrobot", max_steps, seed, render_mode, device, numpy_compat)

    def get_action_lookup(self):
        return {0: "neg torque", 1: "no torque", 2: "pos torque"}

    def get_ob_names(self):
        return ["cos theta1", "sin theta1", "cos theta2", "sin theta2", "dtheta1", "dtheta2", "Gravity",
                "1st Link Length", "2nd Link Length", "1st Link Mass", "2nd Link Mass", "1st Link COM",
                "2nd Link COM", "Moment of Inertia (both links)"]


def create_acrobot(args, hyperparameters, is_valid=False):
    param_range = {
        "gravity": [[9.8, 10.4], [10.4, 24.8]], "link_length_1": [[1., 1.5], [1.5, 2.0]],
        "link_length_2": [[1., 1.5], [1.5, 2.0]], "link_mass_1": [[1., 1.5], [1.5, 2.0]],
        "link_mass_2": [[1., 1.5], [1.5, 2.0]],
    }
    return create_pre_vec(args, hyperparameters, param_range, AcrobotVecEnv, is_valid)
