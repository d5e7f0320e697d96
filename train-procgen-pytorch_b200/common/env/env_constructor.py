"""env_name -> constructor, same names and call signature as the reference (common/env/env_constructor.py:13-31):
``create_venv(args, hyperparameters, is_valid=False)``.  Procgen / MuJoCo names are outside the device hot path and
raise; "lunar_lander" — which the reference refuses — is available here (own semantics, see lunar_lander_pre_vec)."""
from ...boxworld.box_world_env_vec import create_bw_env
from ...discrete_env.acrobot_pre_vec import create_acrobot
from ...discrete_env.cartpole_pre_vec import create_cartpole
from ...discrete_env.cartpole_swing_pre_vec import create_cartpole_swing
from ...discrete_env.lunar_lander_pre_vec import create_lunar_lander
from ...discrete_env.mountain_car_pre_vec import create_mountain_car

_CONSTRUCTORS = {"boxworld": create_bw_env, "cartpole": create_cartpole, "cartpole_swing": create_cartpole_swing,
                 "mountain_car": create_mountain_car, "acrobot": create_acrobot, "lunar_lander": create_lunar_lander}


def get_env_constructor(env_name):
    if env_name in _CONSTRUCTORS:
        return _CONSTRUCTORS[env_name]
    raise NotImplementedError(f"env '{env_name}' is stepped on the host in the reference (Procgen / MuJoCo); "
                              "wrap it as a numpy VecEnv and pass it to PPO, which stages it through Storage.store")
