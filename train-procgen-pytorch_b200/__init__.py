"""B200-native PPO rollout-and-update engine (drop-in for tbuckworth/train-procgen-pytorch's hot path).

Import as ``tpp_b200`` (see /tpp_b200/__init__.py).  Sub-packages mirror the reference's module names:
``discrete_env``, ``boxworld``, ``common`` (storage / model / policy), ``agents`` (ppo).
"""
