"""Cross-check policy engine for the TESTS (not product code): embedder + heads through torch autograd
(cuDNN / cuBLAS), sharing the policy's flat parameter / gradient buffers, with the interface of
tpp_b200.common.engine.ImpalaEngineTC.  Inject with ``PPO(..., engine=TorchModuleEngine(policy, A, obs_shape))``."""
import torch


def _ceil(a, b):
    return (a + b - 1) // b


class TorchModuleEngine:
    """Library path: embedder + heads through torch (cuDNN/cuBLAS) with autograd, sharing the flat buffers."""
    uses_autograd = True      # PPO: no CUDA-graph capture, dhead buffer owned by the agent

    def __init__(self, policy, n_actions, obs_shape):
        assert policy.flat is not None, "call policy.flatten_() first"
        self.policy, self.A = policy, n_actions
        self.ld_head = _ceil(n_actions + 1, 4) * 4
        self.obs_shape = tuple(obs_shape)
        self.device = policy.flat.device
        self.n_launches = 0
        self.last_fs = None

    def forward(self, x, M, feature_major_ld=None, train=False):
        assert feature_major_ld is None
        x = x[:, :int(torch.tensor(self.obs_shape).prod())].reshape(M, *self.obs_shape)
        with torch.set_grad_enabled(train):
            feat, _, fs, _ = self.policy.embedder.forward_with_attn_indices(x)
            logits = self.policy.fc_policy(feat)
            value = self.policy.fc_value(feat)
            head = torch.zeros(M, self.ld_head, dtype=torch.float32, device=self.device)
            head = torch.cat((logits, value, head[:, self.A + 1:]), 1)
        self._head, self.last_fs = head, fs
        return head

    def backward(self, dhead, M, fs_coef=0.0):
        extra = None
        if fs_coef and self.last_fs is not None:
            extra = fs_coef * self.last_fs
        if extra is not None:
            torch.autograd.backward([self._head, extra], [dhead, torch.ones_like(extra)])
        else:
            self._head.backward(dhead)
