"""YAML -> env ctor plumbing (mirrors discrete_env/helper_pre_vec.py:45-62 of the reference).

``defaults`` maps a ctor kwarg to ``[train_value]`` or ``[train_value, validation_value]``; a hyperparameter of
the same name (suffix ``_v`` for the validation env) overrides it; every other non-``_v`` hyperparameter is
passed along and later filtered by the ctor signature.
"""


def assign_env_vars(hyperparameters, is_valid, defaults):
    suffix, pick = ("_v", -1) if is_valid else ("", 0)
    env_args = {}
    for name, values in defaults.items():
        if is_valid and name == "n_envs":
            env_args[name] = hyperparameters.get(name, values[pick])
        env_args[name] = hyperparameters.get(f"{name}{suffix}", values[pick])
    for k, v in hyperparameters.items():
        if k not in env_args and not k.endswith("_v"):
            env_args[k] = v
    return env_args
