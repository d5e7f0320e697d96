"""Agent base class (reference: agents/base_agent.py:3-56): holds env / policy / logger / storage handles."""


class BaseAgent:
    def __init__(self, env, policy, logger, storage, device, num_checkpoints, env_valid=None, storage_valid=None,
                 storage_greedy=None):
        self.env, self.policy, self.logger, self.storage = env, policy, logger, storage
        self.device = device
        self.num_checkpoints = num_checkpoints
        self.env_valid, self.storage_valid, self.storage_greedy = env_valid, storage_valid, storage_greedy
        self.t = 0

    def predict(self, obs):
        raise NotImplementedError

    def optimize(self):
        raise NotImplementedError

    def train(self, num_timesteps):
        raise NotImplementedError
