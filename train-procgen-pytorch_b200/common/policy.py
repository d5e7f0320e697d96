"""CategoricalPolicy: embedder + policy/value heads, same ctor, init, forward contract and state_dict keys as the
reference (common/policy.py:18-87).  ``flatten_()`` re-homes all parameters (and their gradients) into one flat
fp32 buffer laid out for the B200 engine: embedder tensors in state_dict order, then
``fc_policy.weight, fc_value.weight`` (contiguous => one [A+1, D] head matrix) and ``fc_policy.bias,
fc_value.bias`` (one [A+1] head bias)."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.distributions import Categorical

from .model import GRU, orthogonal_init


class CategoricalPolicy(nn.Module):
    def __init__(self, embedder, recurrent, action_size, has_vq=False, continuous_actions=False,
                 logsumexp_logits_is_v=False, extra_params=False):
        super().__init__()
        if has_vq or continuous_actions or logsumexp_logits_is_v or extra_params:
            raise NotImplementedError("VQ / continuous / logsumexp-value / extra-param policies are outside the "
                                      "north-star hot path (SURVEY section 8f)")
        self.embedder = embedder
        self.has_vq, self.continuous_actions, self.recurrent = False, False, bool(recurrent)
        self.action_size = action_size
        self.logsumexp_logits_is_v = False
        self.fc_policy = orthogonal_init(nn.Linear(embedder.output_dim, action_size), gain=0.01)
        self.fc_value = orthogonal_init(nn.Linear(embedder.output_dim, 1), gain=1.0)
        self.target_entropy = np.log(action_size)
        if self.recurrent:      # created last, like the reference (common/policy.py:49-51): same draws, same keys
            self.gru = GRU(embedder.output_dim, embedder.output_dim)
        self.flat = self.flat_grad = None
        self.layout = None

    def is_recurrent(self):
        return self.recurrent

    # ---- reference forward contract (plain torch ops; the engine does not use these) -----------------
    def forward(self, x, hx, masks):
        hidden = self.embedder(x)
        if self.recurrent:
            hidden, hx = self.gru(hidden, hx, masks)
        p, v = self.hidden_to_output(hidden)
        return p, v, hx

    def hidden_to_output(self, hidden):
        logits = self.fc_policy(hidden)
        return self.distribution(logits), self.fc_value(hidden).reshape(-1)

    def distribution(self, logits):
        return Categorical(logits=F.log_softmax(logits, dim=1))

    # ---- flat parameter buffer ------------------------------------------------------------------------
    def flat_order(self):
        names = [f"embedder.{n}" for n, _ in self.embedder.named_parameters()]
        names += ["fc_policy.weight", "fc_value.weight", "fc_policy.bias", "fc_value.bias"]
        if self.recurrent:      # behind the heads: the engines' offsets do not move
            names += [f"gru.gru.{n}" for n in ("weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0")]
        return names

    def flatten_(self, device=None):
        """Move every parameter into one flat fp32 buffer (and .grad into a parallel flat buffer)."""
        params = dict(self.named_parameters())
        order = self.flat_order()
        assert set(order) == set(params), "flat order must cover every parameter exactly once"
        device = torch.device(device) if device is not None else next(self.parameters()).device
        total = sum(params[n].numel() for n in order)
        flat = torch.zeros(total, dtype=torch.float32, device=device)
        gflat = torch.zeros(total, dtype=torch.float32, device=device)
        layout, off = {}, 0
        for n in order:
            p = params[n]
            k = p.numel()
            flat[off:off + k].copy_(p.data.reshape(-1))
            p.data = flat[off:off + k].view(p.shape)
            p.grad = gflat[off:off + k].view(p.shape)
            layout[n] = (off, tuple(p.shape))
            off += k
        self.flat, self.flat_grad, self.layout = flat, gflat, layout
        return self

    def _apply(self, fn, *a, **k):   # .to()/.cuda() after flatten_ would silently break the aliasing
        if self.flat is not None:
            raise RuntimeError("policy was flattened; move it to the device BEFORE flatten_()")
        return super()._apply(fn, *a, **k)
