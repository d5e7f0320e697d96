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


class MLPEngineTC(MLPEngine):
    """MLPModel + heads on the tcgen05 tensor-core GEMM (csrc/gemm_tc.cu).

    Every GEMM operand is a (hi, lo) TF32 pair in its natural row-major layout; forward, data-gradient and
    weight-gradient GEMMs read the same arrays as K-major or MN-major operands, so nothing is transposed.  Pairs are
    written by the producer (GEMM epilogue, minibatch gather, frames_to_obs) or by ``tpp_split_tf32`` (weights after
    an optimizer step, the head's data gradient).  Bias gradients are column sums fused into the epilogue of the
    GEMM that produces the corresponding dZ.  ``precision=3`` (default) is the fp32-parity 3xTF32 mode,
    ``precision=1`` the single-pass fast mode.  The two head GEMMs of the backward pass (A+1 <= 16 columns) stay on
    the CUDA-core kernel: they are < 1 % of the FLOPs.
    """

    def __init__(self, policy, n_actions, precision=3):
        super().__init__(policy, n_actions)
        assert precision in (1, 3)
        self.precision = precision
        f = dict(dtype=torch.float32, device=self.device)
        self.w = []   # per layer: hi, lo [out, ceil32(in)]
        for (w_off, b_off, fin, fout, relu) in self.layers:
            ldk = _ceil(fin, 32) * 32
            self.w.append(dict(hi=torch.zeros(fout, ldk, **f), lo=torch.zeros(fout, ldk, **f), ldk=ldk))
        nh = n_actions + 1
        self.wh = dict(hi=torch.zeros(nh, self.latent, **f), lo=torch.zeros(nh, self.latent, **f))
        self.ld_in = self.w[0]["ldk"]          # row stride the gathered observations must have
        self.refresh_weights()

    # ------------------------------------------------------------------------------------------
    def refresh_weights(self):
        """Re-split the (changed) fp32 weights into TF32 pairs."""
        s = _lib.stream_ptr()
        for (w_off, b_off, fin, fout, relu), w in zip(self.layers, self.w):
            _lib.call("tpp_split_tf32", self._p(w_off), fin, fout, fin, _lib.ptr(w["hi"]), _lib.ptr(w["lo"]), w["ldk"],
                      None, None, 0, s)
        _lib.call("tpp_split_tf32", self._p(self.head_w_off), self.latent, self.A + 1, self.latent,
                  _lib.ptr(self.wh["hi"]), _lib.ptr(self.wh["lo"]), self.latent, None, None, 0, s)
        self.n_launches += len(self.layers) + 1

    def _workspace(self, M):
        ws = self._ws.get(M)
        if ws is None:
            ws = _Workspace()
            f = dict(dtype=torch.float32, device=self.device)
            ws.x = dict(hi=torch.zeros(M, self.ld_in, **f), lo=torch.zeros(M, self.ld_in, **f))
            ws.h = [dict(hi=torch.zeros(M, _ceil(l[3], 32) * 32, **f), lo=torch.zeros(M, _ceil(l[3], 32) * 32, **f),
                         ld=_ceil(l[3], 32) * 32) for l in self.layers]
            ws.last_plain = torch.zeros(M, _ceil(self.latent, 32) * 32, **f)   # same row stride as its TF32 pair
            ws.head = torch.zeros(M, self.ld_head, **f)
            ws.dhead = torch.zeros(M, self.ld_head, **f)
            mw = _ceil(self.max_width, 32) * 32
            ws.dz = [dict(plain=torch.zeros(M, mw, **f), hi=torch.zeros(M, mw, **f), lo=torch.zeros(M, mw, **f))
                     for _ in range(2)]
            ws.fm = None       # TF32 pair of a feature-major rollout slot, allocated on first use
            self._ws[M] = ws
        return ws

    def _tc(self, a, lda, b, ldb, M, N, K, a_mn=0, b_mn=0, flags=0, bias=None, mask=None, ld_mask=0, out=None,
            out_pair=None, ldc=0, colsum=None, split_k=1, block_n=0):
        g = _lib.TcGemm()
        g.a_hi, g.a_lo, g.lda = a[0].data_ptr(), a[1].data_ptr(), lda
        g.b_hi, g.b_lo, g.ldb = b[0].data_ptr(), b[1].data_ptr(), ldb
        g.M, g.N, g.K, g.a_mn, g.b_mn = M, N, K, a_mn, b_mn
        g.precision, g.split_k, g.flags, g.block_n = self.precision, split_k, flags, block_n
        if bias is not None:
            g.bias = bias.value
        if mask is not None:
            g.mask, g.ld_mask = mask.data_ptr(), ld_mask
        if out is not None:
            g.out = out.value if hasattr(out, "value") else out.data_ptr()
        if out_pair is not None:
            g.out_hi, g.out_lo = out_pair[0].data_ptr(), out_pair[1].data_ptr()
        if colsum is not None:
            g.colsum = colsum.value
        g.ldc = ldc
        _lib.call("tpp_gemm_tc", _lib.C.byref(g), _lib.stream_ptr())
        self.n_launches += 1

    # ------------------------------------------------------------------------------------------
    def forward(self, x, M, feature_major_ld=None, x_lo=None, need_backward=True):
        """x: row-major [M, >= in_dim] (plain fp32, or the hi half of a TF32 pair when ``x_lo`` is given), or with
        ``feature_major_ld`` a feature-major [in_dim, ld] rollout slot."""
        ws, s = self._workspace(M), _lib.stream_ptr()
        L = len(self.layers)
        if feature_major_ld:
            ld = feature_major_ld
            if ws.fm is None or ws.fm[0].shape[1] != ld:
                ws.fm = (torch.zeros(self.in_dim, ld, dtype=torch.float32, device=self.device),
                         torch.zeros(self.in_dim, ld, dtype=torch.float32, device=self.device))
            _lib.call("tpp_split_tf32", _lib.ptr(x), ld, self.in_dim, M, _lib.ptr(ws.fm[0]), _lib.ptr(ws.fm[1]), ld,
                      None, None, 0, s)
            self.n_launches += 1
            cur, ld_cur, a_mn = ws.fm, ld, 1
        elif x_lo is not None:
            cur, ld_cur, a_mn = (x, x_lo), x.stride(0), 0
        else:
            _lib.call("tpp_split_tf32", _lib.ptr(x), x.stride(0), M, self.in_dim, _lib.ptr(ws.x["hi"]),
                      _lib.ptr(ws.x["lo"]), self.ld_in, None, None, 0, s)
            self.n_launches += 1
            cur, ld_cur, a_mn = (ws.x["hi"], ws.x["lo"]), self.ld_in, 0
        self._x_pair, self._x_ld = (cur, ld_cur), None
        for i in range(L):
            w_off, b_off, fin, fout, relu = self.layers[i]
            h, w = ws.h[i], self.w[i]
            self._tc(cur, ld_cur, (w["hi"], w["lo"]), w["ldk"], M, fout, fin, a_mn=a_mn,
                     flags=EPI_BIAS | (EPI_RELU if relu else 0), bias=self._p(b_off),
                     out=ws.last_plain if i == L - 1 else None, out_pair=(h["hi"], h["lo"]), ldc=h["ld"])
            cur, ld_cur, a_mn = (h["hi"], h["lo"]), h["ld"], 0
        self._tc(cur, ld_cur, (self.wh["hi"], self.wh["lo"]), self.latent, M, self.A + 1, self.latent, flags=EPI_BIAS,
                 bias=self._p(self.head_b_off), out=ws.head, ldc=self.ld_head, block_n=16)
        self._x = (x, feature_major_ld)
        return ws.head

    def backward(self, dhead, M):
        ws, s = self._workspace(M), _lib.stream_ptr()
        x, fm_ld = self._x
        assert not fm_ld, "backward needs the row-major (minibatch) forward"
        (x_pair, x_ld) = self._x_pair
        H, nh, L = self.latent, self.A + 1, len(self.layers)
        ldl = ws.h[-1]["ld"]
        dz = ws.dz[0]
        last_relu = self.layers[-1][4]
        if H in (16, 32, 64, 128, 256) and nh <= 16:
            # one fused kernel: dlatent (TF32 pair), gWh, gbh and the last embedder layer's bias gradient
            _lib.call("tpp_head_backward", _lib.ptr(dhead), self.ld_head, _lib.ptr(ws.last_plain),
                      _lib.ptr(ws.h[-1]["hi"]) if last_relu else None, ldl, self._p(self.head_w_off), nh, H,
                      _lib.ptr(dz["hi"]), _lib.ptr(dz["lo"]), None, ldl, self._g(self.head_w_off),
                      self._g(self.head_b_off), self._g(self.layers[-1][1]), M, s)
            self.n_launches += 1
        else:
            # generic shapes: heads on the CUDA-core GEMM: gWh += dhead^T latent ; gbh += colsum ; dlatent = dhead Wh
            self._gemm(_lib.ptr(dhead), 1, self.ld_head, _lib.ptr(ws.last_plain), 1, ldl, self._g(self.head_w_off), H,
                       None, None, nh, H, M, EPI_ACCUM, self._split_k(M, nh, H))
            _lib.call("tpp_colsum_accum", _lib.ptr(dhead), self.ld_head, M, nh, self._g(self.head_b_off), s)
            self._gemm(_lib.ptr(dhead), self.ld_head, 1, self._p(self.head_w_off), 1, H, _lib.ptr(dz["plain"]), ldl,
                       None, _lib.ptr(ws.h[-1]["hi"]) if last_relu else None, M, H, nh, EPI_MASK if last_relu else 0)
            _lib.call("tpp_split_tf32", _lib.ptr(dz["plain"]), ldl, M, H, _lib.ptr(dz["hi"]), _lib.ptr(dz["lo"]), ldl,
                      None, None, 0, s)
            _lib.call("tpp_colsum_accum", _lib.ptr(dz["plain"]), ldl, M, H, self._g(self.layers[-1][1]), s)
            self.n_launches += 3
        cur, ld_dz = 0, ldl
        for i in range(L - 1, -1, -1):
            w_off, b_off, fin, fout, relu = self.layers[i]
            dz = ws.dz[cur]
            inp, ld_inp = ((ws.h[i - 1]["hi"], ws.h[i - 1]["lo"]), ws.h[i - 1]["ld"]) if i > 0 else (x_pair, x_ld)
            # gW[fout, fin] += dZ^T X : both operands MN-major, contraction over the M samples split across CTAs
            tiles = _ceil(fout, 128) * _ceil(fin, 128)
            self._tc((dz["hi"], dz["lo"]), ld_dz, inp, ld_inp, fout, fin, M, a_mn=1, b_mn=1, flags=EPI_ACCUM,
                     out=self._g(w_off), ldc=fin, split_k=max(1, min(_ceil(M, 32), _ceil(148, tiles))), block_n=128)
            if i > 0:
                # dZ_{i-1} = (dZ_i W_i) * relu'(H_{i-1}); W_i read as an MN-major operand; column sums = bias grad
                nxt, w, prev = ws.dz[cur ^ 1], self.w[i], ws.h[i - 1]
                prev_relu = self.layers[i - 1][4]
                self._tc((dz["hi"], dz["lo"]), ld_dz, (w["hi"], w["lo"]), w["ldk"], M, fin, fout, b_mn=1,
                         flags=EPI_MASK if prev_relu else 0, mask=prev["hi"] if prev_relu else None,
                         ld_mask=prev["ld"], out_pair=(nxt["hi"], nxt["lo"]), ldc=prev["ld"],
                         colsum=self._g(self.layers[i - 1][1]))
                cur, ld_dz = cur ^ 1, prev["ld"]
