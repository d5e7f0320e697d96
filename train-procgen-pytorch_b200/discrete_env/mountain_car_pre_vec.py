"""mountain_car_pre_vec: 5-column state [position, velocity, gravity, right_boundary, goal_position], start
space rejection-sampled so that goal_position <= right_boundary, sparse reward -1
(reference: discrete_env/mountain_car_pre_vec.py:15-209, 312-328)."""
import numpy as np

from .pre_vec_env import PreVecEnv, create_pre_vec


class MountainCarVecEnv(PreVecEnv):
    family = "mountain_car"
    n_state = 5
    n_obs = 5

    def __init__(self, n_envs=2, goal_velocity=0, left_boundary=-1.2, min_start_position=-0.6,
                 max_start_position=-0.4, max_speed=0.07, min_goal_position=0.5, max_goal_position=3,
                 min_gravity=0.001, max_gravity=0.0025, min_right_boundary=0.6, max_right_boundary=5, force=0.001,
                 max_steps=500, sparse_rewards=True, seed=0, drop_same=False, render_mode=None, device="cuda",
                 numpy_compat=False):
        self.drop_same = drop_same
        assert min_start_position >= left_boundary, \
            f"min_start_position ({min_start_position}) must be >= left_boundary ({left_boundary})"
        assert max_start_position <= min_right_boundary, \
            f"max_start_position ({max_start_position}) must be <= min_right_boundary ({min_right_boundary})"
        assert max_goal_position <= max_right_boundary, \
            f"max_goal_position ({max_goal_position}) must be <= max_right_boundary ({max_right_boundary})"
        for k, v in dict(goal_velocity=goal_velocity, left_boundary=left_boundary,
                         min_start_position=min_start_position, max_start_position=max_start_position,
                         max_speed=max_speed, min_goal_position=min_goal_position,
                         max_goal_position=max_goal_position, min_gravity=min_gravity, max_gravity=max_gravity,
                         min_right_boundary=min_right_boundary, max_right_boundary=max_right_boundary, force=force,
                         sparse_rewards=sparse_rewards).items():
            setattr(self, k, v)
        self.low = np.array([left_boundary, -max_speed, min_gravity, min_right_boundary, min_goal_position],
                            dtype=np.float32)
        self.high = np.array([max_right_boundary, max_speed, max_gravity, max_right_boundary, max_goal_position],
                             dtype=np.float32)
        self.start_low = [min_start_position, 0, min_gravity, min_right_boundary, min_goal_position]
        self.start_high = [max_start_position, 0, max_gravity, max_right_boundary, max_goal_position]
        self.kernel_params = [force, max_speed, left_boundary, goal_velocity, 1.0 if sparse_rewards else 0.0]
        self.customizable_params = ["goal_velocity", "min_start_position", "max_start_position", "left_boundary",
                                    "min_right_boundary", "max_right_boundary", "min_goal_position",
                                    "max_goal_position", "max_speed", "force", "max_steps", "max_gravity",
                                    "min_gravity", "sparse_rewards"]
        super().__init__(n_envs, 3, "MountainCar", max_steps, seed, render_mode, device, numpy_compat)

    def get_ob_names(self):
        return ["Position", "Velocity", "Gravity", "Right Boundary", "Goal Position"]

    def get_action_lookup(self):
        return {0: "acc left", 1: "none", 2: "acc right"}


def create_mountain_car(args, hyperparameters, is_valid=False):
    param_range = {
        "goal_velocity": [0], "left_boundary": [-1.2], "min_right_boundary": [0.6, 1.],
        "max_right_boundary": [1., 5.], "max_speed": [0.07], "min_goal_position": [0, 0.6],
        "max_goal_position": [0.6, 3], "force": [0.001], "min_gravity": [0.001, 0.0015],
        "max_gravity": [0.0015, 0.0025], "sparse_rewards": [True],
    }
    return create_pre_vec(args, hyperparameters, param_range, MountainCarVecEnv, is_valid)
