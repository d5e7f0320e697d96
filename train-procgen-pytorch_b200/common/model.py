"""Policy embedders: MLPModel and ImpalaModel with the reference's architecture, init and state_dict keys
(reference: common/model.py:134-208 IMPALA, :954-980 MLP; inits common/misc_util.py:78-89).

Both are ``torch.nn.Module``s so that checkpoints interchange with the reference
(``embedder.model.*`` / ``embedder.block{1,2,3}.*`` / ``embedder.fc.*``).  After ``CategoricalPolicy.flatten_()``
every parameter is a view into ONE flat fp32 buffer; the B200 engine (common/engine.py) runs its own CUDA
kernels on that buffer, while the ``nn.Module.forward`` defined here (plain torch ops) is kept for
API compatibility and as the in-test autograd cross-check.
"""
from __future__ import annotations

import torch
import torch.nn as nn


def xavier_uniform_init(module, gain=1.0):
    if isinstance(module, (nn.Linear, nn.Conv2d)):
        nn.init.xavier_uniform_(module.weight.data, gain)
        nn.init.constant_(module.bias.data, 0)
    return module


def orthogonal_init(module, gain=nn.init.calculate_gain("relu")):
    if isinstance(module, (nn.Linear, nn.Conv2d)):
        nn.init.orthogonal_(module.weight.data, gain)
        nn.init.constant_(module.bias.data, 0)
    return module


class MLPModel(nn.Module):
    """Linear(in, mid)-ReLU-[Linear(mid, mid)-ReLU] x (depth-2)-Linear(mid, latent); no final activation."""

    def __init__(self, in_channels, depth, mid_weight, latent_size, normalize=False):
        super().__init__()
        if normalize:
            raise NotImplementedError("LayerNorm variant is not used by any reference config on the hot path")
        self.input_size, self.depth, self.mid_weight, self.output_dim = in_channels, depth, mid_weight, latent_size
        mid = []
        for _ in range(depth - 2):
            mid += [nn.Linear(mid_weight, mid_weight), nn.ReLU()]
        self.model = nn.Sequential(nn.Linear(in_channels, mid_weight), nn.ReLU(), nn.Sequential(*mid),
                                   nn.Linear(mid_weight, latent_size))
        self.apply(xavier_uniform_init)

    def dense_layers(self):
        """[(Linear module, relu_after)] in forward order — consumed by the engine."""
        lins = [m for m in self.model.modules() if isinstance(m, nn.Linear)]
        return [(lin, i < len(lins) - 1) for i, lin in enumerate(lins)]

    def forward(self, x):
        return self.model(x)

    def forward_with_attn_indices(self, x):
        return self.model(x), [], None, None


class ResidualBlock(nn.Module):
    def __init__(self, in_channels):
        super().__init__()
        self.conv1 = nn.Conv2d(in_channels, in_channels, kernel_size=3, stride=1, padding=1)
        self.conv2 = nn.Conv2d(in_channels, in_channels, kernel_size=3, stride=1, padding=1)

    def forward(self, x):
        out = self.conv1(torch.relu(x))
        out = self.conv2(torch.relu(out))
        return out + x


class ImpalaBlock(nn.Module):
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size=3, stride=1, padding=1)
        self.res1 = ResidualBlock(out_channels)
        self.res2 = ResidualBlock(out_channels)

    def forward(self, x):
        x = nn.functional.max_pool2d(self.conv(x), kernel_size=3, stride=2, padding=1)
        return self.res2(self.res1(x))


class ImpalaModel(nn.Module):
    """3 IMPALA blocks (16, 32, latent_dim channels) -> ReLU -> flatten -> FC -> ReLU.

    ``input_hw``: the reference hard-codes an 8x8 final map (64x64 frames, common/model.py:175); passing the
    frame size makes the same architecture usable on Box-World's 14x14 frames."""

    def __init__(self, in_channels, output_dim=256, latent_dim=32, input_hw=(64, 64), **kwargs):
        super().__init__()
        self.block1 = ImpalaBlock(in_channels, 16)
        self.block2 = ImpalaBlock(16, 32)
        self.block3 = ImpalaBlock(32, latent_dim)
        h, w = input_hw
        for _ in range(3):
            h, w = (h + 1) // 2, (w + 1) // 2
        self.encoded_dim = latent_dim * h * w
        self.fc = nn.Linear(self.encoded_dim, output_dim)
        self.output_dim = output_dim
        self.apply(xavier_uniform_init)

    def forward_to_pool(self, x):
        x = torch.relu(self.block3(self.block2(self.block1(x))))
        return x.flatten(1)

    def forward_from_pool(self, h):
        return torch.relu(self.fc(h))

    def forward(self, x):
        return self.forward_from_pool(self.forward_to_pool(x))

    def forward_with_attn_indices(self, x):
        h = self.forward_to_pool(x)
        out = self.forward_from_pool(h)
        feature_sparsity = torch.mean(torch.max(torch.tanh(torch.abs(h * 100)), 0)[0])
        return out, [], feature_sparsity, None


class GRU(nn.Module):
    """The reference's recurrent core (common/model.py:212-276): ``nn.GRU(input, hidden)`` under the name ``gru`` -- the
    reference passes it through ``orthogonal_init``, which only touches Linear / Conv2d, so it keeps torch's default
    uniform initialisation (and its draws from the global generator).  ``forward(x, hxs, masks)``: with one row per
    hidden state (prediction) a single cell step on ``hxs * masks``; otherwise ``x`` is a ``(T, N, .)`` batch flattened
    to ``(T*N, .)`` and the hidden state is re-computed through time, reset where ``masks`` is zero.  This module is the
    torch statement of the contract (tests, checkpoints); rollouts run ``tpp_gru_cell`` on its parameters."""

    def __init__(self, input_size, hidden_size):
        super().__init__()
        self.gru = orthogonal_init(nn.GRU(input_size, hidden_size), gain=1.0)

    def forward(self, x, hxs, masks):
        if x.size(0) == hxs.size(0):
            x, hxs = self.gru(x.unsqueeze(0), (hxs * masks.unsqueeze(-1)).unsqueeze(0))
            return x.squeeze(0), hxs.squeeze(0)
        N = hxs.size(0)
        T = x.size(0) // N
        x, masks = x.view(T, N, x.size(1)), masks.view(T, N)
        h, outs = hxs, []
        for t in range(T):          # the reference batches runs of mask == 1 into one cuDNN call: the same recurrence
            h = self.gru(x[t:t + 1], (h * masks[t].unsqueeze(-1)).unsqueeze(0))[1].squeeze(0)
            outs.append(h)
        return torch.stack(outs).view(T * N, -1), h
