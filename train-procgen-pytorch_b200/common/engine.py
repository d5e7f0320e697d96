"""Policy compute engines: run the policy forward/backward on the flat parameter buffer with this repo's CUDA
kernels (no autograd on the hot path).

* ``MLPEngine``   — MLPModel + heads: every layer is one C-ABI GEMM call with fused bias/ReLU(-mask) epilogue;
                    backward accumulates straight into the flat gradient buffer.
* ``TorchModuleEngine`` — library path (cuDNN/cuBLAS through torch autograd) for embedders that do not yet have
                    hand-written kernels (IMPALA convolutions this round).  The heads' output still feeds the
                    fused loss kernel and the flat-buffer Adam; only the embedder's contraction is library code.

Replaces: ``policy.embedder.forward_with_attn_indices`` + ``policy.hidden_to_output`` + ``loss.backward()``
(agents/ppo.py:125-128,170) and ``policy(obs, hx, mask)`` in ``PPO.predict`` (agents/ppo.py:76).
"""
from __future__ import annotations

import torch

from .. import _lib
from .._lib import EPI_ACCUM, EPI_BIAS, EPI_MASK, EPI_RELU


def _ceil(a, b):
    return (a + b - 1) // b


class _Workspace:
    def __init__(self):
        self.acts = []      # activations per layer [M, out]
        self.head = None    # [M, ld_head]
        self.dhead = None
        self.dbuf = None    # two ping-pong gradient buffers [M, max_width]


class MLPEngine:
    def __init__(self, policy, n_actions):
        assert policy.flat is not None, "call policy.flatten_() first"
        self.policy = policy
        self.flat, self.gflat = policy.flat, policy.flat_grad
        self.device = self.flat.device
        self.A = n_actions
        self.ld_head = _ceil(n_actions + 1, 4) * 4
        self.layers = []    # (w_off, b_off, fan_in, fan_out, relu)
        names = {id(p): n for n, p in policy.named_parameters()}
        for lin, relu in policy.embedder.dense_layers():
            w_off = policy.layout[names[id(lin.weight)]][0]
            b_off = policy.layout[names[id(lin.bias)]][0]
            self.layers.append((w_off, b_off, lin.in_features, lin.out_features, relu))
        self.in_dim = self.layers[0][2]
        self.latent = self.layers[-1][3]
        self.head_w_off = policy.layout["fc_policy.weight"][0]
        self.head_b_off = policy.layout["fc_policy.bias"][0]
        assert policy.layout["fc_value.weight"][0] == self.head_w_off + n_actions * self.latent
        assert policy.layout["fc_value.bias"][0] == self.head_b_off + n_actions
        self.max_width = max([l[3] for l in self.layers] + [self.in_dim])
        self._ws = {}
        self.n_launches = 0   # running count of this engine's kernel launches (for bench.py's gpu_launches)

    # ------------------------------------------------------------------------------------------
    def _workspace(self, M):
        ws = self._ws.get(M)
        if ws is None:
            ws = _Workspace()
            dev = self.device
            ws.acts = [torch.empty(M, l[3], dtype=torch.float32, device=dev) for l in self.layers]
            ws.head = torch.zeros(M, self.ld_head, dtype=torch.float32, device=dev)
            ws.dhead = torch.zeros(M, self.ld_head, dtype=torch.float32, device=dev)
            ws.dbuf = [torch.empty(M, self.max_width, dtype=torch.float32, device=dev) for _ in range(2)]
            self._ws[M] = ws
        return ws

    def _p(self, off):
        return _lib.C.c_void_p(self.flat.data_ptr() + 4 * off)

    def _g(self, off):
        return _lib.C.c_void_p(self.gflat.data_ptr() + 4 * off)

    def _gemm(self, A, sam, sak, B, sbn, sbk, Cc, ldc, bias, mask, M, N, K, flags, split_k=1):
        _lib.call("tpp_gemm_f32", A, sam, sak, B, sbn, sbk, Cc, ldc, bias, mask, M, N, K, flags, split_k,
                  _lib.stream_ptr())
        self.n_launches += 1

    # ------------------------------------------------------------------------------------------
    def forward(self, x, M, feature_major_ld=None):
        """x: row-major [M, in_dim] tensor, or (with ``feature_major_ld``) a feature-major [in_dim, ld] rollout
        slot.  Returns the head buffer [M, ld_head] = (A logits, value, zero padding)."""
        ws = self._workspace(M)
        a_ptr = _lib.ptr(x)
        sam, sak = (1, feature_major_ld) if feature_major_ld else (x.stride(0), 1)
        for i, (w_off, b_off, fin, fout, relu) in enumerate(self.layers):
            out = ws.acts[i]
            self._gemm(a_ptr, sam, sak, self._p(w_off), fin, 1, _lib.ptr(out), fout, self._p(b_off), None, M, fout,
                       fin, EPI_BIAS | (EPI_RELU if relu else 0))
            a_ptr, sam, sak = _lib.ptr(out), fout, 1
        self._gemm(a_ptr, sam, sak, self._p(self.head_w_off), self.latent, 1, _lib.ptr(ws.head), self.ld_head,
                   self._p(self.head_b_off), None, M, self.A + 1, self.latent, EPI_BIAS)
        self._x = (x, feature_major_ld)
        return ws.head

    def backward(self, dhead, M):
        """Accumulate d loss / d params into the flat gradient buffer given d loss / d head [M, ld_head].
        Must follow ``forward`` on the same M (activations are read from the workspace)."""
        ws = self._workspace(M)
        x, fm_ld = self._x
        H, nh = self.latent, self.A + 1
        split = self._split_k(M, nh, H)
        # heads: gW += dhead^T @ latent ; gb += colsum(dhead) ; dlatent = dhead @ Wh
        last = ws.acts[-1]
        self._gemm(_lib.ptr(dhead), 1, self.ld_head, _lib.ptr(last), 1, H, self._g(self.head_w_off), H, None, None,
                   nh, H, M, EPI_ACCUM, split)
        _lib.call("tpp_colsum_accum", _lib.ptr(dhead), self.ld_head, M, nh, self._g(self.head_b_off),
                  _lib.stream_ptr())
        self.n_launches += 1
        dz, cur = ws.dbuf[0], 0
        self._gemm(_lib.ptr(dhead), self.ld_head, 1, self._p(self.head_w_off), 1, H, _lib.ptr(dz), H, None,
                   _lib.ptr(last) if self.layers[-1][4] else None, M, H, nh,
                   EPI_MASK if self.layers[-1][4] else 0)
        for i in range(len(self.layers) - 1, -1, -1):
            w_off, b_off, fin, fout, relu = self.layers[i]
            # dz: [M, fout] gradient w.r.t. this layer's pre-activation
            if i > 0:
                inp_ptr, isam, isak = _lib.ptr(ws.acts[i - 1]), self.layers[i - 1][3], 1
            elif fm_ld:
                inp_ptr, isam, isak = _lib.ptr(x), 1, fm_ld
            else:
                inp_ptr, isam, isak = _lib.ptr(x), x.stride(0), 1
            # gW[fout, fin] += dz^T @ inp   (contraction over the M samples, split across CTAs)
            self._gemm(_lib.ptr(dz), 1, fout, inp_ptr, isak, isam, self._g(w_off), fin, None, None, fout, fin, M,
                       EPI_ACCUM, self._split_k(M, fout, fin))
            _lib.call("tpp_colsum_accum", _lib.ptr(dz), fout, M, fout, self._g(b_off), _lib.stream_ptr())
            self.n_launches += 1
            if i > 0:
                # d(prev pre-activation) = (dz @ W) * relu'(prev activation)
                nxt = ws.dbuf[cur ^ 1]
                prev_relu = self.layers[i - 1][4]
                self._gemm(_lib.ptr(dz), fout, 1, self._p(w_off), 1, fin, _lib.ptr(nxt), fin, None,
                           _lib.ptr(ws.acts[i - 1]) if prev_relu else None, M, fin, fout,
                           EPI_MASK if prev_relu else 0)
                dz, cur = nxt, cur ^ 1

    @staticmethod
    def _split_k(M, rows, cols):
        tiles = _ceil(rows, 64) * _ceil(cols, 64)
        return max(1, min(_ceil(M, 256), _ceil(296, tiles)))


class TorchModuleEngine:
    """Library path: embedder + heads through torch (cuDNN/cuBLAS) with autograd, sharing the flat buffers."""

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
