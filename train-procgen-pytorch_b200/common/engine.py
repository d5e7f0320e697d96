"""Policy compute engines: run the policy forward/backward on the flat parameter buffer with this repo's CUDA
kernels (no autograd on the hot path).

* ``MLPEngine``   — MLPModel + heads: every layer is one C-ABI GEMM call with fused bias/ReLU(-mask) epilogue;
                    backward accumulates straight into the flat gradient buffer.
* ``MLPEngineTC`` — the same on the tcgen05 tensor-core GEMM (3xTF32 parity mode or single-pass TF32).
* ``ImpalaEngineTC`` — ImpalaModel + heads: NHWC activations, 3x3 convolutions as im2col + tcgen05 GEMM with the
                    residual adds / ReLU masks / bias gradients fused into the GEMM epilogues, max-pool kernels.
(The cuDNN / cuBLAS cross-check engine used by the tests lives in tests/torch_engine.py: it is not part of the product.)

Replaces: ``policy.embedder.forward_with_attn_indices`` + ``policy.hidden_to_output`` + ``loss.backward()``
(agents/ppo.py:125-128,170) and ``policy(obs, hx, mask)`` in ``PPO.predict`` (agents/ppo.py:76).
"""
from __future__ import annotations

import os

import torch

from .. import _lib
from .._lib import TC_TILE_PAIR64_PERSISTENT, TC_TILE_PAIR_PERSISTENT, TC_TILE_PAIR_PERSISTENT_LEAN, TC_A_EXACT, TC_B_EXACT, TC_A_SPLIT, TC_B_SPLIT, EPI_ACCUM, EPI_ADD, EPI_BIAS, EPI_MASK, EPI_PAIR_RELU, EPI_RELU, EPI_RELU_OUT


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
    def _workspace(self, M, slot=0):
        ws = self._ws.get((M, slot))
        if ws is None:
            ws = _Workspace()
            dev = self.device
            ws.acts = [torch.empty(M, l[3], dtype=torch.float32, device=dev) for l in self.layers]
            ws.head = torch.zeros(M, self.ld_head, dtype=torch.float32, device=dev)
            ws.dhead = torch.zeros(M, self.ld_head, dtype=torch.float32, device=dev)
            ws.dbuf = [torch.empty(M, self.max_width, dtype=torch.float32, device=dev) for _ in range(2)]
            self._ws[(M, slot)] = ws
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
    def forward(self, x, M, feature_major_ld=None, slot=0):
        """x: row-major [M, in_dim] tensor, or (with ``feature_major_ld``) a feature-major [in_dim, ld] rollout
        slot.  Returns the head buffer [M, ld_head] = (A logits, value, zero padding).  ``slot`` selects a private
        workspace (concurrent forwards of different env ranges on different streams)."""
        ws = self._workspace(M, slot)
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

    def __init__(self, policy, n_actions, precision=3, raw_pixels=False, obs_shape=None, split_on_chip=False):
        super().__init__(policy, n_actions)
        assert precision in (1, 3)
        self.precision = precision
        # split_on_chip: hidden activations and their gradients live in HBM as ONE plain fp32 array; the GEMMs that read
        # them form the lo halves in shared memory (TPP_TC_A_SPLIT / TPP_TC_B_SPLIT), so every large operand is read
        # once and every epilogue writes once.  Needs the wide tiles (> 32 output columns) in every GEMM of the trunk.
        self.split_on_chip = bool(split_on_chip) and precision == 3 and \
            all(l[3] > 32 for l in self.layers) and all(l[2] > 32 for l in self.layers[1:])
        # raw_pixels: image observations may arrive as integer pixel values 0..255 (``forward(..., raw=True)``).  They are
        # exact in TF32, so the first layer needs no lo half of X (two MMA passes, 3/4 of the operand traffic, half the
        # gather's writes); ScaledFloatFrame's 1/255 (common/env/procgen_wrappers.py:407-419) lives in a second copy of
        # the first layer's weights and in the ``alpha`` of its weight-gradient GEMM.
        self.raw_pixels = raw_pixels
        # 256 x 256 output tiles on CTA pairs (cta_group::2) for the 256-wide layers once a pass has enough rows to fill
        # the machine with them: each CTA stages 128 rows of each operand for a 256 x 256 product, i.e. half the
        # L2 -> shared-memory bytes per FLOP of 128 x 128 tiles, which is what bounds these kernels (profiles/README.md)
        self.wide_tile_rows = 32768
        self.pair_min_n = 256
        self.pair_block_n = TC_TILE_PAIR_PERSISTENT     # or TC_TILE_PAIR: one 256 x 256 tile per (non-persistent) pair
        # launches whose epilogue 8 warps can hide (see TPP_TC_TILE_PAIR_PERSISTENT_LEAN): three operand stages
        self.lean_kinds = ("fwd", "dgrad", "wgrad")
        # contractions shorter than this keep the 16-epilogue-warp tile: two k-blocks of main loop cannot hide a lean epilogue
        # (the 256 <- 64 data gradient: 89.6 us lean, 73.7 us with 16 epilogue warps at 131072 rows)
        self.lean_min_k = int(os.environ.get("TPP_LEAN_MIN_K", "128"))
        self.small_tile_elems = 148 * 128 * 128 // 2
        f = dict(dtype=torch.float32, device=self.device)
        self.w = []   # per layer: hi, lo [out, ceil32(in)]
        for (w_off, b_off, fin, fout, relu) in self.layers:
            ldk = _ceil(fin, 32) * 32
            self.w.append(dict(hi=torch.zeros(fout, ldk, **f), lo=torch.zeros(fout, ldk, **f), ldk=ldk))
        nh = n_actions + 1
        self.wh = dict(hi=torch.zeros(nh, self.latent, **f), lo=torch.zeros(nh, self.latent, **f))
        self.ld_in = self.w[0]["ldk"]          # row stride the gathered observations must have
        if raw_pixels:
            fin, fout, ldk = self.layers[0][2], self.layers[0][3], self.w[0]["ldk"]
            self.w0_scaled = torch.zeros(fout, fin, **f)
            self.w0_raw = dict(hi=torch.zeros(fout, ldk, **f), lo=torch.zeros(fout, ldk, **f), ldk=ldk)
            # the same weights with their columns in FRAME BYTE ORDER (NHWC: byte p*C + c <-> feature c*H*W + p): the
            # fused rollout kernel then reads a uint8 frame as its first-layer operand without TransposeFrame
            self.w0_bytes = None
            if obs_shape is not None and len(obs_shape) == 3 and int(obs_shape[0] * obs_shape[1] * obs_shape[2]) == fin:
                c, h, w = (int(v) for v in obs_shape)
                b = torch.arange(fin, device=self.device)
                self._byte_perm = (b % c) * (h * w) + b // c
                self.w0_bytes_plain = torch.zeros(fout, fin, **f)
                self.w0_bytes = dict(hi=torch.zeros(fout, ldk, **f), lo=torch.zeros(fout, ldk, **f), ldk=ldk)
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
        if self.raw_pixels:
            w_off, _, fin, fout, _ = self.layers[0]
            torch.mul(self.flat[w_off:w_off + fout * fin].view(fout, fin), 1.0 / 255.0, out=self.w0_scaled)
            _lib.call("tpp_split_tf32", _lib.ptr(self.w0_scaled), fin, fout, fin, _lib.ptr(self.w0_raw["hi"]),
                      _lib.ptr(self.w0_raw["lo"]), self.w0_raw["ldk"], None, None, 0, s)
            self.n_launches += 1
            if self.w0_bytes is not None:
                torch.index_select(self.w0_scaled, 1, self._byte_perm, out=self.w0_bytes_plain)
                _lib.call("tpp_split_tf32", _lib.ptr(self.w0_bytes_plain), fin, fout, fin, _lib.ptr(self.w0_bytes["hi"]),
                          _lib.ptr(self.w0_bytes["lo"]), self.w0_bytes["ldk"], None, None, 0, s)
                self.n_launches += 1

    def weight_views(self):
        """(ctypes array of tpp_weight_view, count): where the TF32 operand copies of every parameter live, for
        ``tpp_adam_clip_step_views`` -- the optimizer step then leaves them current and ``refresh_weights`` is not
        needed behind it."""
        if getattr(self, "_views", None) is None:
            views = []

            def add(off, rows, cols, w, scale=0.0, col_of=None):
                v = _lib.WeightView()
                v.offset, v.rows, v.cols, v.ld, v.scale = off, rows, cols, w["ldk"] if "ldk" in w else cols, scale
                v.hi, v.lo = w["hi"].data_ptr(), w["lo"].data_ptr()
                v.col_of = col_of.data_ptr() if col_of is not None else None
                views.append(v)
            for (w_off, b_off, fin, fout, relu), w in zip(self.layers, self.w):
                add(w_off, fout, fin, w)
            add(self.head_w_off, self.A + 1, self.latent, self.wh)
            if self.raw_pixels:
                w_off, _, fin, fout, _ = self.layers[0]
                add(w_off, fout, fin, self.w0_raw, 1.0 / 255.0)
                if self.w0_bytes is not None:
                    inv = torch.empty_like(self._byte_perm)
                    inv[self._byte_perm] = torch.arange(fin, device=self.device)
                    self._byte_col = inv.to(torch.int32)       # feature c of a weight row sits at frame byte inv[c]
                    add(w_off, fout, fin, self.w0_bytes, 1.0 / 255.0, self._byte_col)
            assert len(views) <= _lib.MAX_WEIGHT_VIEWS
            self._views = ((_lib.WeightView * len(views))(*views), len(views))
        return self._views

    def _workspace(self, M, slot=0):
        ws = self._ws.get((M, slot))
        if ws is None:
            ws = _Workspace()
            f = dict(dtype=torch.float32, device=self.device)
            ws.x = dict(hi=torch.zeros(M, self.ld_in, **f), lo=torch.zeros(M, self.ld_in, **f))
            L = len(self.layers)
            ws.h = []
            for i, l in enumerate(self.layers):
                hi = torch.zeros(M, _ceil(l[3], 32) * 32, **f)
                # split_on_chip: "hi" is the plain activation, only the head GEMM's operand (last layer) keeps a lo half
                lo = hi if (self.split_on_chip and i < L - 1) else torch.zeros(M, _ceil(l[3], 32) * 32, **f)
                ws.h.append(dict(hi=hi, lo=lo, ld=_ceil(l[3], 32) * 32))
            ws.last_plain = torch.zeros(M, _ceil(self.latent, 32) * 32, **f)   # same row stride as its TF32 pair
            ws.head = torch.zeros(M, self.ld_head, **f)
            ws.dhead = torch.zeros(M, self.ld_head, **f)
            mw = _ceil(self.max_width, 32) * 32
            ws.dz = [dict(plain=torch.zeros(M, mw, **f), hi=torch.zeros(M, mw, **f), lo=torch.zeros(M, mw, **f))
                     for _ in range(2)]     # split_on_chip: "hi" of ws.dz[1] / later ws.dz[0] is the plain gradient
            ws.fm = None       # TF32 pair of a feature-major rollout slot, allocated on first use
            # 1-bit ReLU masks of the hidden layers (written by the forward epilogue, read by the data gradient): 1/32 of
            # the fp32 activation the mask would otherwise be re-read from
            ws.bits = None
            if M % 32 == 0 and all(l[3] % 32 == 0 and l[3] > 32 for l in self.layers[:-1]) and self.precision in (1, 3):
                ws.bits = [torch.zeros(M // 32 * (l[3] // 32) * 32, dtype=torch.int32, device=self.device)
                           for l in self.layers[:-1]]
            self._ws[(M, slot)] = ws
        return ws

    def _tc(self, a, lda, b, ldb, M, N, K, a_mn=0, b_mn=0, flags=0, bias=None, mask=None, ld_mask=0, out=None,
            out_pair=None, ldc=0, colsum=None, split_k=1, block_n=0, addend=None, ld_add=0, conv=None, conv_wgrad=0,
            exact=0, alpha=0.0, mask_bits_out=None, mask_bits=None):
        g = _lib.TcGemm()
        g.a_hi, g.a_lo, g.lda = a[0].data_ptr(), a[1].data_ptr(), lda
        g.b_hi, g.b_lo, g.ldb = b[0].data_ptr(), b[1].data_ptr(), ldb
        g.M, g.N, g.K, g.a_mn, g.b_mn = M, N, K, a_mn, b_mn
        g.precision, g.split_k, g.flags, g.block_n = self.precision | (exact if self.precision == 3 else 0), split_k, \
            flags, block_n
        g.alpha = alpha
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
        if mask_bits_out is not None:
            g.mask_bits_out = mask_bits_out.data_ptr()
        if mask_bits is not None:
            g.mask_bits = mask_bits.data_ptr()
        if addend is not None:
            g.addend, g.ld_add = addend.data_ptr(), ld_add
        if conv is not None:
            g.conv_B, g.conv_H, g.conv_W, g.conv_C = conv
            g.conv_wgrad = conv_wgrad
        g.ldc = ldc
        _lib.call("tpp_gemm_tc", _lib.C.byref(g), _lib.stream_ptr())
        self.n_launches += 1

    # ------------------------------------------------------------------------------------------
    def tail_ok(self):
        """The rollout tail kernel covers the last embedder layer + heads (shape limits of tpp_mlp_tail_sample)."""
        if len(self.layers) < 2:
            return False
        _, _, fin, fout, _ = self.layers[-1]
        return fin <= 256 and fin % 4 == 0 and fout == 64 and self.A + 1 <= 16

    def fused_rollout_ok(self, raw=False):
        """Shapes ``tpp_policy_rollout_fused`` is built for: the reference's depth-4 ``mlpmodel`` sets (hidden 256, latent
        64) in the fp32-parity 3xTF32 mode; the first layer either reads raw pixel rows (<= 768 inputs) or a feature-major
        vector slot (<= 128 inputs)."""
        if self.precision != 3 or len(self.layers) != 4 or self.A + 1 > 16:
            return False
        if [l[3] for l in self.layers] != [256, 256, 256, 64] or [l[2] for l in self.layers[1:]] != [256, 256, 256]:
            return False
        return self.in_dim <= (768 if raw else 128)

    def rollout_fused(self, x, M, ldx, raw, act, logp, value, seed, tick, t_offset, env_offset=0, greedy=False,
                      head_out=None, dbg=None):
        """One launch: whole forward + heads + action draw for M rows (csrc/rollout_fused.cu).  ``raw``: True = x is
        row-major integer pixel rows [M][ldx] fp32 (exact TF32 operand, 1/255 folded into the first layer's weight
        copy); "u8" = x is the uint8 NHWC frames [M][ldx bytes] themselves (weight columns in frame byte order);
        False = x is a feature-major rollout slot [in_dim][ldx]."""
        f = _lib.FusedPolicy()
        f.n_rows, f.a1_mode, f.x, f.ldx = M, {True: 0, False: 1, "u8": 2}[raw], x.data_ptr(), ldx
        for i, (w_off, b_off, fin, fout, relu) in enumerate(self.layers):
            w = (self.w0_bytes if raw == "u8" else self.w0_raw) if (raw and i == 0) else self.w[i]
            f.w_hi[i], f.w_lo[i], f.ldw[i] = w["hi"].data_ptr(), w["lo"].data_ptr(), w["ldk"]
            f.k[i], f.n[i], f.relu[i] = fin, fout, 1 if relu else 0
            f.bias[i] = self.flat.data_ptr() + 4 * b_off
        f.head_w = self.flat.data_ptr() + 4 * self.head_w_off
        f.head_b = self.flat.data_ptr() + 4 * self.head_b_off
        f.n_actions, f.ld_head = self.A, self.ld_head
        f.act, f.logp, f.value = act.data_ptr(), logp.data_ptr(), value.data_ptr()
        f.head_out = head_out.data_ptr() if head_out is not None else None
        f.seed, f.tick, f.t_offset = seed, tick.data_ptr(), int(t_offset)
        f.greedy, f.env_offset = 1 if greedy else 0, int(env_offset)
        f.dbg = dbg.data_ptr() if dbg is not None else None
        if getattr(self, "_fused_scratch", None) is None:      # exchange slots of up to 37 concurrent clusters (148 SMs / 4)
            self._fused_scratch = torch.zeros(37 * 4 * 3 * 128 * 256, dtype=torch.uint8, device=self.device)
        f.scratch, f.scratch_bytes = self._fused_scratch.data_ptr(), self._fused_scratch.numel()
        _lib.call("tpp_policy_rollout_fused", _lib.C.byref(f), _lib.stream_ptr())
        self.n_launches += 1

    def _bn(self, M, N, kind="fwd", K=None):
        """Tile of a forward / data-gradient GEMM with M rows and N output columns (tpp_tc_gemm.block_n)."""
        if N >= self.pair_min_n and M >= self.wide_tile_rows:
            lean = kind in self.lean_kinds and not (K is not None and K < self.lean_min_k)
            if self.pair_block_n == TC_TILE_PAIR_PERSISTENT and lean and self.precision == 3:
                return TC_TILE_PAIR_PERSISTENT_LEAN
            return self.pair_block_n       # 256 x 256 tiles on CTA pairs (cta_group::2)
        if 32 < N <= 64 and M >= self.wide_tile_rows and self.pair_block_n == TC_TILE_PAIR_PERSISTENT:
            return TC_TILE_PAIR64_PERSISTENT   # 256 x 64 tiles on persistent CTA pairs (HBM-bound: A is read once)
        if N >= 128 and M * N <= self.small_tile_elems:
            return 64          # few tiles (rollout forward, M = n_envs): 128 x 64 tiles put twice as many SMs to work
        return 0

    def forward(self, x, M, feature_major_ld=None, x_lo=None, need_backward=True, raw=False, slot=0, trunk_only=False,
                heads=True):
        """x: row-major [M, >= in_dim] (plain fp32, or the hi half of a TF32 pair when ``x_lo`` is given, or -- ``raw`` --
        integer pixel values 0..255 with row stride ``ld_in``), or with ``feature_major_ld`` a feature-major
        [in_dim, ld] rollout slot.  ``slot``: private workspace for concurrent forwards on different streams.
        ``trunk_only``: stop in front of the last embedder layer and return its input as plain fp32 ``(tensor, ld)`` --
        the rollout finishes that layer, the heads and the action draw in one CUDA-core launch
        (``tpp_mlp_tail_sample``).  ``heads=False``: stop behind the embedder and return its output as the TF32 pair
        ``((hi, lo), ld)`` -- a recurrent policy puts its GRU cell between the embedder and ``head_gemm``."""
        ws, s = self._workspace(M, slot), _lib.stream_ptr()
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
        elif raw:
            assert self.raw_pixels
            cur, ld_cur, a_mn = (x, x), x.stride(0), 0      # exact operand: the lo half is never loaded
        elif x_lo is not None:
            cur, ld_cur, a_mn = (x, x_lo), x.stride(0), 0
        else:
            _lib.call("tpp_split_tf32", _lib.ptr(x), x.stride(0), M, self.in_dim, _lib.ptr(ws.x["hi"]),
                      _lib.ptr(ws.x["lo"]), self.ld_in, None, None, 0, s)
            self.n_launches += 1
            cur, ld_cur, a_mn = (ws.x["hi"], ws.x["lo"]), self.ld_in, 0
        self._x_pair, self._x_ld, self._x_raw = (cur, ld_cur), None, raw
        self._bits_valid = bool(need_backward and ws.bits is not None and not trunk_only and not feature_major_ld)
        for i in range(L - 1 if trunk_only else L):
            w_off, b_off, fin, fout, relu = self.layers[i]
            h, w = ws.h[i], (self.w0_raw if raw and i == 0 else self.w[i])
            soc = self.split_on_chip
            a_flag = TC_A_EXACT if raw and i == 0 else (TC_A_SPLIT if soc and i > 0 else 0)
            if trunk_only and i == L - 2:       # plain fp32 only: its consumer is not a tensor-core GEMM
                self._tc(cur, ld_cur, (w["hi"], w["lo"]), w["ldk"], M, fout, fin, a_mn=a_mn,
                         flags=EPI_BIAS | (EPI_RELU if relu else 0), bias=self._p(b_off), out=h["hi"], ldc=h["ld"],
                         exact=a_flag, block_n=self._bn(M, fout))
                return h["hi"], h["ld"]
            plain_only = soc and i < L - 1     # the next GEMM splits it on chip
            bits = ws.bits[i] if (ws.bits is not None and need_backward and relu and i < L - 1) else None
            self._tc(cur, ld_cur, (w["hi"], w["lo"]), w["ldk"], M, fout, fin, a_mn=a_mn, mask_bits_out=bits,
                     flags=EPI_BIAS | (EPI_RELU if relu else 0), bias=self._p(b_off),
                     out=ws.last_plain if i == L - 1 else (h["hi"] if plain_only else None),
                     out_pair=None if plain_only else (h["hi"], h["lo"]), ldc=h["ld"],
                     exact=a_flag, block_n=self._bn(M, fout))
            cur, ld_cur, a_mn = (h["hi"], h["lo"]), h["ld"], 0
        self._x = (x, feature_major_ld)
        if not heads:
            return cur, ld_cur
        return self.head_gemm(cur, ld_cur, M, slot)

    def head_gemm(self, latent_pair, ld, M, slot=0):
        """[A logits | value] = latent Wh^T + bh for M rows of a TF32 latent pair -> the workspace's head buffer."""
        ws = self._workspace(M, slot)
        self._tc(latent_pair, ld, (self.wh["hi"], self.wh["lo"]), self.latent, M, self.A + 1, self.latent, flags=EPI_BIAS,
                 bias=self._p(self.head_b_off), out=ws.head, ldc=self.ld_head, block_n=16)
        return ws.head

    @staticmethod
    def _wgrad_split(M, ctas):
        """k-splits of a weight-gradient GEMM (contraction over the M samples; ``ctas`` CTAs per split): at most ~1024
        samples per TMEM accumulator (the tensor core adds into it with truncation; longer chains lose ~1e-4,
        DESIGN.md 'Long contractions'), at least one wave of CTAs, and then as many splits as fit the same number of
        waves (150 CTAs on 148 SMs take as long as 296)."""
        s_min = max(_ceil(148, ctas), _ceil(M, 1024))
        waves = _ceil(s_min * ctas, 148)
        return max(1, min(_ceil(M, 32), max(s_min, (waves * 148) // ctas)))

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
            pair = fout >= 256 and fin >= 256 and M >= self.wide_tile_rows      # 256 x 256 tiles on CTA pairs
            ctas = 2 * _ceil(fout, 256) * _ceil(fin, 256) if pair else _ceil(fout, 128) * _ceil(fin, 128)
            raw0 = i == 0 and self._x_raw         # gW1 = (1/255) dZ^T X_pixels, X exact: no lo half, two passes
            soc = self.split_on_chip
            # split_on_chip: dZ of layers below the last and the hidden activations are plain arrays (the head kernel's
            # dZ of the last layer and the gathered observations are pairs)
            flag = (TC_B_EXACT if raw0 else (TC_B_SPLIT if soc and i > 0 else 0)) | (TC_A_SPLIT if soc and i < L - 1 else 0)
            self._tc((dz["hi"], dz["lo"]), ld_dz, inp, ld_inp, fout, fin, M, a_mn=1, b_mn=1, flags=EPI_ACCUM,
                     out=self._g(w_off), ldc=fin, split_k=self._wgrad_split(M, ctas),
                     block_n=(TC_TILE_PAIR_PERSISTENT_LEAN if "wgrad" in self.lean_kinds and self.precision == 3 and
                              self.pair_block_n == TC_TILE_PAIR_PERSISTENT else self.pair_block_n) if pair else 128,
                     exact=flag, alpha=1.0 / 255.0 if raw0 else 0.0)
            if i > 0:
                # dZ_{i-1} = (dZ_i W_i) * relu'(H_{i-1}); W_i read as an MN-major operand; column sums = bias grad
                nxt, w, prev = ws.dz[cur ^ 1], self.w[i], ws.h[i - 1]
                prev_relu = self.layers[i - 1][4]
                use_bits = prev_relu and ws.bits is not None and self._bits_valid
                self._tc((dz["hi"], dz["lo"]), ld_dz, (w["hi"], w["lo"]), w["ldk"], M, fin, fout, b_mn=1,
                         mask_bits=ws.bits[i - 1] if use_bits else None,
                         flags=EPI_MASK if (prev_relu and not use_bits) else 0,
                         mask=prev["hi"] if (prev_relu and not use_bits) else None,
                         ld_mask=prev["ld"], out=nxt["hi"] if soc else None,
                         out_pair=None if soc else (nxt["hi"], nxt["lo"]), ldc=prev["ld"],
                         colsum=self._g(self.layers[i - 1][1]), block_n=self._bn(M, fin, "dgrad", K=fout),
                         exact=TC_A_SPLIT if soc and i < L - 1 else 0)
                cur, ld_dz = cur ^ 1, prev["ld"]


class ImpalaEngineTC:
    """ImpalaModel (common/model.py:81-114; reference common/model.py:134-208) + heads on this repo's kernels.

    Activations are NHWC fp32 ``[B*H*W, C]`` matrices, kept both plain (ReLU masks, skip connections) and as the TF32
    (hi, lo) pair the tensor cores read.  A 3x3 / pad-1 convolution is ONE ``tpp_gemm_tc`` launch in convolution mode:
    the A tiles are gathered from the NHWC pair by TMA im2col loads (implicit GEMM, no col matrix), the bias, the
    residual add (``TPP_EPI_ADD``), the ReLU in front of the next convolution (``TPP_EPI_PAIR_RELU``: the pair output
    is ReLU'd, the plain output is not) and -- behind the last block -- the trailing ReLU are fused into the epilogue.
    The data gradient is the same launch on the dY pair with tap-flipped weights, with the ReLU mask of the
    convolution's input, the skip-path gradient and the bias gradient of the layer below (column sums) fused into the
    epilogue.  The weight gradient is ``dY^T col(X)`` with both operands MN-major and the contraction over B*H*W split
    across CTAs; its col matrix is still materialised (``tpp_im2col3x3``), as is the first convolution's (3 input
    channels).  Weight copies in GEMM layout are rebuilt from the flat fp32 parameters by ``refresh_weights`` after
    every optimizer step; weight gradients are accumulated in GEMM layout and folded back into the flat gradient buffer
    at the end of ``backward``.

    ``x`` for ``forward`` is what ``tpp_gather_img`` / ``tpp_frames_to_obs`` produce: fp32 rows ``[M][ld]`` holding
    the frame as NCHW / 255 (the first im2col reads it through strides, nothing is transposed).
    The feature-sparsity term of ``forward_with_attn_indices`` (common/model.py:203-208), fs = mean_j max_b
    tanh(|100 h_bj|), is computed by ``tpp_feature_sparsity`` on the ReLU'd block-3 features (value logged every
    minibatch like the reference); with ``fs_coef != 0`` its gradient is added to the feature gradient in front of
    block 3's backward pass (``tpp_feature_sparsity_grad``).
    """

    _tc = MLPEngineTC._tc
    WGRAD_CHUNK = 1024   # rows contracted per CTA before the partial sums meet in IEEE fp32 atomics
    KI = 288             # contraction length of an implicit convolution: 9 taps x 32 channel slots

    def __init__(self, policy, n_actions, obs_shape, precision=3):
        assert policy.flat is not None, "call policy.flatten_() first"
        assert precision in (1, 3)
        self.policy, self.A, self.precision = policy, n_actions, precision
        self.flat, self.gflat = policy.flat, policy.flat_grad
        self.device = self.flat.device
        self.ld_head = _ceil(n_actions + 1, 4) * 4
        self.obs_shape = tuple(obs_shape)
        self.n_launches = 0
        self.last_fs = None
        self._ws = {}
        # weight gradients of the 16 / 32-channel convolutions on the fp32 FMA pipe (tpp_conv3x3_wgrad: exact fp32, each
        # operand byte staged once; 3-4x the tensor-core form at these channel counts); False = tcgen05 im2col form
        self.wgrad_cc = os.environ.get("TPP_WGRAD_FMA", "1") != "0"
        # forward / data gradient of the convolutions with a 16-channel side at 32 x 32 on the FMA pipe too (tpp_conv3x3_fma)
        # -- opt-in (TPP_FMA_FWD=1): measured 303 / 336 us against 335 / 349 us of the tensor-core tiles at 16 -> 16 channels,
        # but slower at 16 <-> 32, and no gain over a whole forward + backward (11.19 vs 11.03 ms per 2048 frames)
        self.fma_fwd = os.environ.get("TPP_FMA_FWD", "0") == "1"
        emb = policy.embedder
        self._names = {id(p): n for n, p in policy.named_parameters()}
        C, H, W = self.obs_shape
        self.convs, self.blocks = [], []
        h, w, cin = H, W, C
        for blk in (emb.block1, emb.block2, emb.block3):
            cout = blk.conv.out_channels
            assert cout in (16, 32), "conv widths must be 16 or 32 (GEMM N tile, 32 channel slots per tap)"
            ci = self._add_conv(blk.conv, cin, cout, need_dgrad=len(self.convs) > 0)
            ho, wo = (h + 1) // 2, (w + 1) // 2
            res = [(self._add_conv(rb.conv1, cout, cout), self._add_conv(rb.conv2, cout, cout))
                   for rb in (blk.res1, blk.res2)]
            self.blocks.append(dict(conv=ci, res=res, H=h, W=w, Ho=ho, Wo=wo, cin=cin, cout=cout))
            h, w, cin = ho, wo, cout
        self.enc_hw, self.enc_c, self.enc = h * w, cin, h * w * cin
        assert emb.fc.in_features == self.enc and self.enc % 32 == 0
        self.latent = emb.fc.out_features
        self.fc_w_off, self.fc_b_off = self._off(emb.fc.weight), self._off(emb.fc.bias)
        self.head_w_off = policy.layout["fc_policy.weight"][0]
        self.head_b_off = policy.layout["fc_policy.bias"][0]
        assert policy.layout["fc_value.weight"][0] == self.head_w_off + n_actions * self.latent
        assert policy.layout["fc_value.bias"][0] == self.head_b_off + n_actions
        f = dict(dtype=torch.float32, device=self.device)
        # weight-gradient accumulators in GEMM layout: one zeroable buffer, one view per layer
        # (implicit convolutions: [tap*32 + channel slot][cout]; explicit first convolution: [tap*cin + ci][cout])
        sizes = [c["cout"] * (self.KI if c["implicit"] else c["Kf"]) for c in self.convs] + [self.latent * self.enc]
        self.gtmp = torch.zeros(sum(sizes), **f)
        o = 0
        for c, n in zip(self.convs, sizes[:-1]):
            c["gw"] = self.gtmp[o:o + n].view(self.KI if c["implicit"] else c["Kf"], c["cout"])
            o += n
        self.gfc = self.gtmp[o:o + sizes[-1]].view(self.latent, self.enc)
        self.wfc_plain = torch.zeros(self.latent, self.enc, **f)
        self.wfc = (torch.zeros(self.latent, self.enc, **f), torch.zeros(self.latent, self.enc, **f))
        nh = n_actions + 1
        self.wh = (torch.zeros(nh, self.latent, **f), torch.zeros(nh, self.latent, **f))
        self.refresh_weights()

    # ------------------------------------------------------------------------------------------
    def _off(self, param):
        return self.policy.layout[self._names[id(param)]][0]

    def _p(self, off):
        return _lib.C.c_void_p(self.flat.data_ptr() + 4 * off)

    def _g(self, off):
        return _lib.C.c_void_p(self.gflat.data_ptr() + 4 * off)

    def _add_conv(self, conv, cin, cout, need_dgrad=True):
        f = dict(dtype=torch.float32, device=self.device)
        Kf = _ceil(9 * cin, 32) * 32                     # explicit col width (weight gradient; forward of conv 1)
        implicit = cin % 4 == 0 and cin <= 32            # forward through TMA im2col
        # channel slots per tap of the implicit forms: 16 (64-byte rows) when the gathered tensor has <= 16 channels
        sf, sd = (16 if cin <= 16 else 32), (16 if cout <= 16 else 32)
        kw = 9 * sf if implicit else Kf
        c = dict(w_off=self._off(conv.weight), b_off=self._off(conv.bias), cin=cin, cout=cout, Kf=Kf, implicit=implicit,
                 sf=sf, sd=sd,
                 wf_plain=torch.zeros(cout, kw, **f), wf=(torch.zeros(cout, kw, **f), torch.zeros(cout, kw, **f)))
        if need_dgrad:
            c["wd_plain"] = torch.zeros(cin, 9 * sd, **f)
            c["wd"] = (torch.zeros(cin, 9 * sd, **f), torch.zeros(cin, 9 * sd, **f))
        self.convs.append(c)
        return len(self.convs) - 1

    def _split(self, plain, pair):
        rows, cols = plain.shape
        _lib.call("tpp_split_tf32", _lib.ptr(plain), cols, rows, cols, _lib.ptr(pair[0]), _lib.ptr(pair[1]), cols,
                  None, None, 0, _lib.stream_ptr())
        self.n_launches += 1

    def refresh_weights(self):
        """Rebuild the GEMM-layout TF32 weight copies from the flat fp32 parameters (after an optimizer step)."""
        for c in self.convs:
            cin, cout = c["cin"], c["cout"]
            w = self.flat[c["w_off"]:c["w_off"] + cout * cin * 9].view(cout, cin, 3, 3)
            if c["implicit"]:    # Wf[co][tap][32 slots]
                c["wf_plain"].view(cout, 3, 3, c["sf"])[..., :cin].copy_(w.permute(0, 2, 3, 1))
            else:                # Wf[co][tap*cin + ci]
                c["wf_plain"][:, :9 * cin].view(cout, 3, 3, cin).copy_(w.permute(0, 2, 3, 1))
            self._split(c["wf_plain"], c["wf"])
            if "wd" in c:        # Wd[ci][tap][32 slots] = W[co][ci][2-ky][2-kx]
                c["wd_plain"].view(cin, 3, 3, c["sd"])[..., :cout].copy_(w.flip(2, 3).permute(1, 2, 3, 0))
                self._split(c["wd_plain"], c["wd"])
        wfc = self.flat[self.fc_w_off:self.fc_w_off + self.latent * self.enc].view(self.latent, self.enc_c, self.enc_hw)
        self.wfc_plain.view(self.latent, self.enc_hw, self.enc_c).copy_(wfc.permute(0, 2, 1))
        self._split(self.wfc_plain, self.wfc)
        wh = self.flat[self.head_w_off:self.head_w_off + (self.A + 1) * self.latent].view(self.A + 1, self.latent)
        self._split(wh, self.wh)

    # ------------------------------------------------------------------------------------------
    def _workspace(self, M):
        ws = self._ws.get(M)
        if ws is not None:
            return ws
        ws = _Workspace()
        f = dict(dtype=torch.float32, device=self.device)

        def trio(n):
            return dict(plain=torch.zeros(n, **f), hi=torch.zeros(n, **f), lo=torch.zeros(n, **f))

        col = 0
        ws.blk = []
        for b in self.blocks:
            rows_in, rows = M * b["H"] * b["W"], M * b["Ho"] * b["Wo"]
            cout = b["cout"]
            if not self.convs[b["conv"]]["implicit"]:
                col = max(col, rows_in * self.convs[b["conv"]]["Kf"])
            ws.blk.append(dict(a=torch.zeros(rows_in * cout, **f),
                               arg=torch.zeros(rows * cout, dtype=torch.uint8, device=self.device),
                               p=trio(rows * cout), c1=trio(rows * cout), r1=trio(rows * cout), c2=trio(rows * cout),
                               r2=trio(rows * cout),
                               gX=trio(rows * cout), gY=trio(rows * cout), gZ=trio(rows * cout),
                               ga=trio(rows_in * cout)))
        ws.col = (torch.zeros(max(col, 4), **f), torch.zeros(max(col, 4), **f))   # col matrix of the first convolution
        ws.col_src = None
        ws.h = (torch.zeros(M, self.enc, **f), torch.zeros(M, self.enc, **f))        # relu(block3), NHWC-flattened
        ws.f = (torch.zeros(M, self.latent, **f), torch.zeros(M, self.latent, **f))  # relu(fc)
        ws.last_plain = torch.zeros(M, self.latent, **f)
        ws.fc_acc = torch.zeros(M, self.latent, **f)
        ws.dz = (torch.zeros(M, self.latent, **f), torch.zeros(M, self.latent, **f))
        ws.head = torch.zeros(M, self.ld_head, **f)
        ws.dhead = torch.zeros(M, self.ld_head, **f)
        ws.fs_key = torch.zeros(self.enc, dtype=torch.int64, device=self.device)   # (max, argmax row) per feature
        ws.fs = torch.zeros(1, **f)
        self._ws[M] = ws
        return ws

    def _im2col(self, src, B, H, W, C, strides, relu, col, Kp):
        _lib.call("tpp_im2col3x3", _lib.ptr(src), 0, B, H, W, C, strides[0], strides[1], strides[2], strides[3],
                  1 if relu else 0, 1.0, _lib.ptr(col[0]), _lib.ptr(col[1]) if self.precision == 3 else None, Kp,
                  _lib.stream_ptr())
        self.n_launches += 1

    @staticmethod
    def _nhwc(H, W, C):
        return (H * W * C, W * C, C, 1)

    def _conv_fwd(self, ws, ci, src, B, H, W, out=None, addend=None, relu_out=False, pair_relu=False, pair=None,
                  strides=None, plain_in=None, relu_in=True):
        """src: TF32 pair of the (already ReLU'd) NHWC input for an implicit convolution, or -- first convolution --
        the plain source tensor with its element strides.  out: trio (plain + pair written) or plain tensor."""
        c = self.convs[ci]
        rows = B * H * W
        flags = EPI_BIAS | (EPI_ADD if addend is not None else 0) | (EPI_RELU_OUT if relu_out else 0) | \
            (EPI_PAIR_RELU if pair_relu else 0)
        plain = out["plain"] if isinstance(out, dict) else out
        if pair is None and isinstance(out, dict):
            pair = (out["hi"], out["lo"])
        kw = dict(flags=flags, bias=self._p(c["b_off"]), out=plain, out_pair=pair, ldc=c["cout"], addend=addend,
                  ld_add=c["cout"])
        if c["implicit"] and plain_in is not None and self.fma_fwd and not relu_out:
            # a 16-channel tensor on either side at 32 x 32: exact-fp32 FMA kernel (tpp_conv3x3_fma); other shapes answer
            # ENOTSUP and take the tensor-core tile below
            if _lib.try_call("tpp_conv3x3_fma", _lib.ptr(plain_in), 1 if relu_in else 0, _lib.ptr(c["wf_plain"]), c["sf"],
                             self._p(c["b_off"]), None, _lib.ptr(addend), 1 if pair_relu else 0, _lib.ptr(plain),
                             _lib.ptr(pair[0]) if pair else None, _lib.ptr(pair[1]) if pair else None, None, B, H, W,
                             c["cin"], c["cout"], _lib.stream_ptr()):
                self.n_launches += 1
                return
        if c["implicit"]:
            self._tc(src, 0, c["wf"], 9 * c["sf"], rows, c["cout"], 9 * c["sf"], conv=(B, H, W, c["cin"]), **kw)
        elif self._first_cc(c, H, W, strides) and flags == EPI_BIAS and pair is None:
            # first convolution (K = 27): FMA-pipe kernel straight from the planar observation and the flat fp32
            # parameters -- no col matrix, no TF32 weight copies
            _lib.call("tpp_conv3x3_fwd_first", _lib.ptr(src), strides[0], strides[3], strides[1], self._p(c["w_off"]),
                      self._p(c["b_off"]), _lib.ptr(plain), B, H, W, c["cout"], _lib.stream_ptr())
            self.n_launches += 1
        else:
            self._im2col(src, B, H, W, c["cin"], strides, False, ws.col, c["Kf"])
            ws.col_src = src
            self._tc(ws.col, c["Kf"], c["wf"], c["Kf"], rows, c["cout"], c["Kf"], **kw)

    def forward(self, x, M, feature_major_ld=None, need_backward=True, train=False, heads=True):
        assert feature_major_ld is None
        ws = self._workspace(M)
        C0, H0, W0 = self.obs_shape
        nb = len(self.blocks)
        for k, (b, wb) in enumerate(zip(self.blocks, ws.blk)):
            H, W, Ho, Wo, cout = b["H"], b["W"], b["Ho"], b["Wo"], b["cout"]
            if k == 0:
                self._conv_fwd(ws, b["conv"], x, M, H, W, out=wb["a"], strides=(x.stride(0), W0, 1, H0 * W0))
            else:
                prev = ws.blk[k - 1]["r2"]
                self._conv_fwd(ws, b["conv"], (prev["hi"], prev["lo"]), M, H, W, out=wb["a"], plain_in=prev["plain"],
                               relu_in=False)
            p, c1, r1, c2, r2 = wb["p"], wb["c1"], wb["r1"], wb["c2"], wb["r2"]
            _lib.call("tpp_maxpool3x3s2_fwd", _lib.ptr(wb["a"]), M, H, W, cout, _lib.ptr(p["plain"]), _lib.ptr(wb["arg"]),
                      _lib.ptr(p["hi"]), _lib.ptr(p["lo"]), _lib.stream_ptr())
            self.n_launches += 1
            (a1, b1), (a2, b2) = b["res"]
            self._conv_fwd(ws, a1, (p["hi"], p["lo"]), M, Ho, Wo, out=c1, pair_relu=True, plain_in=p["plain"])
            self._conv_fwd(ws, b1, (c1["hi"], c1["lo"]), M, Ho, Wo, out=r1, addend=p["plain"], pair_relu=True,
                           plain_in=c1["plain"])
            self._conv_fwd(ws, a2, (r1["hi"], r1["lo"]), M, Ho, Wo, out=c2, pair_relu=True, plain_in=r1["plain"])
            if k == nb - 1:   # trailing ReLU + flatten: written directly as the fc layer's TF32 operand
                self._conv_fwd(ws, b2, (c2["hi"], c2["lo"]), M, Ho, Wo, pair=ws.h, addend=r1["plain"], relu_out=True)
            else:             # the next block's convolution reads r2 without a ReLU
                self._conv_fwd(ws, b2, (c2["hi"], c2["lo"]), M, Ho, Wo, out=r2, addend=r1["plain"], plain_in=c2["plain"])
        if self.enc > 512 and self.precision == 3:
            # long contraction: chunks of 256 accumulated with IEEE adds (the tensor core truncates when it adds into
            # its fp32 accumulator), then bias + ReLU + TF32 split
            ws.fc_acc.zero_()
            self._tc(ws.h, self.enc, self.wfc, self.enc, M, self.latent, self.enc, flags=EPI_ACCUM, out=ws.fc_acc,
                     ldc=self.latent, split_k=_ceil(self.enc, 256))
            _lib.call("tpp_bias_act_split", _lib.ptr(ws.fc_acc), self.latent, M, self.latent, self._p(self.fc_b_off), 1,
                      _lib.ptr(ws.last_plain), _lib.ptr(ws.f[0]), _lib.ptr(ws.f[1]), self.latent, _lib.stream_ptr())
            self.n_launches += 1
        else:
            self._tc(ws.h, self.enc, self.wfc, self.enc, M, self.latent, self.enc, flags=EPI_BIAS | EPI_RELU,
                     bias=self._p(self.fc_b_off), out=ws.last_plain, out_pair=ws.f, ldc=self.latent)
        self._x = x
        if not heads:            # recurrent policies: the GRU cell sits between the fc layer and head_gemm
            return ws.f, self.latent
        self.head_gemm(ws.f, self.latent, M)
        if train:   # feature sparsity (logged every minibatch; enters the loss when fs_coef != 0)
            _lib.call("tpp_feature_sparsity", _lib.ptr(ws.h[0]), _lib.ptr(ws.h[1]), M, self.enc, _lib.ptr(ws.fs_key),
                      _lib.ptr(ws.fs), _lib.stream_ptr())
            self.n_launches += 2
            self.last_fs = ws.fs[0]
        return ws.head

    def head_gemm(self, latent_pair, ld, M, slot=0):
        ws = self._workspace(M)
        self._tc(latent_pair, ld, self.wh, self.latent, M, self.A + 1, self.latent, flags=EPI_BIAS,
                 bias=self._p(self.head_b_off), out=ws.head, ldc=self.ld_head, block_n=16)
        return ws.head

    # ------------------------------------------------------------------------------------------
    def _wgrad(self, ws, ci, dy, src, B, H, W, strides=None, plain=None, relu=True):
        """Weight gradient from the dY pair and the convolution's input.  Implicit form (src = TF32 pair of the NHWC
        input, already ReLU'd): gw[tap*32 + ci][co] += sum_p X[p + tap][ci] dY[p][co], the A tiles gathered by TMA im2col.
        Explicit form (first convolution, src = plain tensor + strides): gw[co][tap*cin + ci] += dY^T col(X)."""
        c = self.convs[ci]
        rows = B * H * W
        chunks = _ceil(rows, self.WGRAD_CHUNK)
        # (32 -> 32 channels at 16 x 16: the tensor-core form is the faster one, 241 vs 257 us at 2048 frames -- measured)
        if c["implicit"] and plain is not None and self.wgrad_cc and (c["cin"], c["cout"], H) != (32, 32, 16):
            # narrow layers: exact-fp32 FMA kernel with the halo staged once per tile (csrc/conv_cc.cu); shapes it was
            # not built for answer ENOTSUP and take the tensor-core form below
            if _lib.try_call("tpp_conv3x3_wgrad", _lib.ptr(plain), 1 if relu else 0, _lib.ptr(dy["plain"]),
                             _lib.ptr(c["gw"]), B, H, W, c["cin"], c["cout"], _lib.stream_ptr()):
                self.n_launches += 1
                return
        if c["implicit"]:
            self._tc(src, 0, (dy["hi"], dy["lo"]), c["cout"], self.KI, c["cout"], rows, a_mn=1, b_mn=1, flags=EPI_ACCUM,
                     out=c["gw"], ldc=c["cout"], block_n=32, conv=(B, H, W, c["cin"]), conv_wgrad=1,
                     split_k=max(1, min(_ceil(rows, 32), max(_ceil(296, 3), chunks))))
            return
        if self._first_cc(c, H, W, strides):
            # first convolution (3 -> 16 on the planar 64 x 64 observation): FMA-pipe kernel straight from the observation
            # and the plain dY (no col matrix, no dY pair)
            _lib.call("tpp_conv3x3_wgrad_first", _lib.ptr(src), strides[0], strides[3], strides[1], _lib.ptr(dy["plain"]),
                      _lib.ptr(c["gw"]), B, H, W, c["cout"], _lib.stream_ptr())
            self.n_launches += 1
            return
        if ws.col_src is not src:        # the forward pass of this minibatch left col(X) in place
            self._im2col(src, B, H, W, c["cin"], strides, False, ws.col, c["Kf"])
            ws.col_src = src
        n = 9 * c["cin"]
        # transposed product gw[k][co] = col^T dY: the wide operand fills the 128-row M tile, N = cout stays narrow
        # (small CTAs, several per SM)
        self._tc(ws.col, c["Kf"], (dy["hi"], dy["lo"]), c["cout"], n, c["cout"], rows, a_mn=1, b_mn=1, flags=EPI_ACCUM,
                 out=c["gw"], ldc=c["cout"], block_n=32,
                 split_k=max(1, min(_ceil(rows, 32), max(_ceil(296, _ceil(n, 128)), chunks))))

    def _first_cc(self, c, H, W, strides):
        """Shapes tpp_conv3x3_wgrad_first was built for (everything else: col matrix + tensor-core GEMM)."""
        return bool(self.wgrad_cc and not c["implicit"] and c["cin"] == 3 and c["cout"] == 16 and H == 64 and W == 64
                    and strides is not None and strides[2] == 1)

    def _dgrad(self, ws, ci, dy, B, H, W, out, mask=None, addend=None, colsum_off=None):
        """dX = conv(dY, flipped W) (* relu mask of the conv input) (+ skip gradient); column sums -> bias grad below."""
        c = self.convs[ci]
        rows = B * H * W
        flags = (EPI_MASK if mask is not None else 0) | (EPI_ADD if addend is not None else 0)
        if self.fma_fwd:
            if _lib.try_call("tpp_conv3x3_fma", _lib.ptr(dy["plain"]), 0, _lib.ptr(c["wd_plain"]), c["sd"], None,
                             _lib.ptr(mask), _lib.ptr(addend), 0, _lib.ptr(out["plain"]), _lib.ptr(out["hi"]),
                             _lib.ptr(out["lo"]), self._g(colsum_off) if colsum_off is not None else None, B, H, W,
                             c["cout"], c["cin"], _lib.stream_ptr()):
                self.n_launches += 1
                return
        self._tc((dy["hi"], dy["lo"]), 0, c["wd"], 9 * c["sd"], rows, c["cin"], 9 * c["sd"], conv=(B, H, W, c["cout"]),
                 flags=flags, mask=mask, ld_mask=c["cin"], out=out["plain"], out_pair=(out["hi"], out["lo"]),
                 ldc=c["cin"], addend=addend, ld_add=c["cin"],
                 colsum=self._g(colsum_off) if colsum_off is not None else None)

    def _colsum(self, x, rows, C, off):
        _lib.call("tpp_colsum_narrow", _lib.ptr(x), rows, C, self._g(off), _lib.stream_ptr())
        self.n_launches += 1

    def backward(self, dhead, M, fs_coef=0.0):
        ws, s = self._workspace(M), _lib.stream_ptr()
        H, nh = self.latent, self.A + 1
        assert H in (16, 32, 64, 128, 256) and nh <= 16
        self.gtmp.zero_()

        def pair(t):
            return (t["hi"], t["lo"])

        _lib.call("tpp_head_backward", _lib.ptr(dhead), self.ld_head, _lib.ptr(ws.last_plain), _lib.ptr(ws.f[0]), H,
                  self._p(self.head_w_off), nh, H, _lib.ptr(ws.dz[0]), _lib.ptr(ws.dz[1]), None, H,
                  self._g(self.head_w_off), self._g(self.head_b_off), self._g(self.fc_b_off), M, s)
        self.n_launches += 1
        # fc: weight gradient in NHWC column order, data gradient masked by the trailing ReLU of block 3
        tiles = _ceil(H, 128) * _ceil(self.enc, 128)
        self._tc(ws.dz, H, ws.h, self.enc, H, self.enc, M, a_mn=1, b_mn=1, flags=EPI_ACCUM, out=self.gfc, ldc=self.enc,
                 split_k=max(1, min(_ceil(M, 32), max(_ceil(148, tiles), _ceil(M, 512)))), block_n=128)
        d = ws.blk[-1]["gX"]
        self._tc(ws.dz, H, self.wfc, self.enc, M, self.enc, H, b_mn=1, flags=EPI_MASK, mask=ws.h[0], ld_mask=self.enc,
                 out=d["plain"], out_pair=(d["hi"], d["lo"]), ldc=self.enc)
        if fs_coef:      # + fs_coef * d feature_sparsity / d h (agents/ppo.py:164-169), needs forward(train=True)
            _lib.call("tpp_feature_sparsity_grad", _lib.ptr(ws.fs_key), self.enc, float(fs_coef), _lib.ptr(d["plain"]),
                      _lib.ptr(d["hi"]), _lib.ptr(d["lo"]), s)
            self.n_launches += 1
        for k in range(len(self.blocks) - 1, -1, -1):
            b, wb = self.blocks[k], ws.blk[k]
            Hh, Ww, Ho, Wo, cout = b["H"], b["W"], b["Ho"], b["Wo"], b["cout"]
            rows = M * Ho * Wo
            st = self._nhwc(Ho, Wo, cout)
            (a1, b1), (a2, b2) = b["res"]
            X, Y, Z = wb["gX"], wb["gY"], wb["gZ"]
            p, c1, r1, c2 = wb["p"]["plain"], wb["c1"]["plain"], wb["r1"]["plain"], wb["c2"]["plain"]
            if k == len(self.blocks) - 1:
                self._colsum(X["plain"], rows, cout, self.convs[b2]["b_off"])
            # res2: r2 = conv_b2(relu(c2)) + r1 ; c2 = conv_a2(relu(r1))
            self._wgrad(ws, b2, X, pair(wb["c2"]), M, Ho, Wo, plain=c2)
            self._dgrad(ws, b2, X, M, Ho, Wo, Y, mask=c2, colsum_off=self.convs[a2]["b_off"])
            self._wgrad(ws, a2, Y, pair(wb["r1"]), M, Ho, Wo, plain=r1)
            self._dgrad(ws, a2, Y, M, Ho, Wo, Z, mask=r1, addend=X["plain"], colsum_off=self.convs[b1]["b_off"])
            # res1: r1 = conv_b1(relu(c1)) + p ; c1 = conv_a1(relu(p))
            self._wgrad(ws, b1, Z, pair(wb["c1"]), M, Ho, Wo, plain=c1)
            self._dgrad(ws, b1, Z, M, Ho, Wo, Y, mask=c1, colsum_off=self.convs[a1]["b_off"])
            self._wgrad(ws, a1, Y, pair(wb["p"]), M, Ho, Wo, plain=p)
            self._dgrad(ws, a1, Y, M, Ho, Wo, X, mask=p, addend=Z["plain"])
            # max-pool, then the block's first convolution
            ga = wb["ga"]
            x_strides = (self._x.stride(0), self.obs_shape[2], 1, self.obs_shape[1] * self.obs_shape[2])
            # block 1: the pooled gradient feeds only the first convolution's weight gradient; when that runs on the FMA
            # kernel it reads the plain tensor and the TF32 pair (2/3 of this kernel's writes) is not produced
            need_pair = not (k == 0 and self._first_cc(self.convs[b["conv"]], Hh, Ww, x_strides))
            _lib.call("tpp_maxpool3x3s2_bwd", _lib.ptr(X["plain"]), _lib.ptr(wb["arg"]), M, Hh, Ww, cout,
                      _lib.ptr(ga["plain"]), _lib.ptr(ga["hi"]) if need_pair else None,
                      _lib.ptr(ga["lo"]) if need_pair else None, s)
            self.n_launches += 1
            self._colsum(ga["plain"], M * Hh * Ww, cout, self.convs[b["conv"]]["b_off"])
            if k == 0:
                self._wgrad(ws, b["conv"], ga, self._x, M, Hh, Ww, strides=x_strides)
            else:
                prev = ws.blk[k - 1]
                self._wgrad(ws, b["conv"], ga, pair(prev["r2"]), M, Hh, Ww, plain=prev["r2"]["plain"], relu=False)
                self._dgrad(ws, b["conv"], ga, M, Hh, Ww, prev["gX"],
                            colsum_off=self.convs[self.blocks[k - 1]["res"][1][1]]["b_off"])
        # fold the GEMM-layout weight gradients back into the flat gradient buffer
        for c in self.convs:
            cin, cout = c["cin"], c["cout"]
            g = self.gflat[c["w_off"]:c["w_off"] + cout * cin * 9].view(cout, cin, 3, 3)
            if c["implicit"]:
                g += c["gw"].view(3, 3, 32, cout)[:, :, :cin, :].permute(3, 2, 0, 1)
            else:
                g += c["gw"][:9 * cin].view(3, 3, cin, cout).permute(3, 2, 0, 1)
        g = self.gflat[self.fc_w_off:self.fc_w_off + self.latent * self.enc].view(self.latent, self.enc_c, self.enc_hw)
        g += self.gfc.view(self.latent, self.enc_hw, self.enc_c).permute(0, 2, 1)


class GRUCellTC:
    """The recurrent core of ``CategoricalPolicy(recurrent=True)`` at prediction time: one ``nn.GRU(D, D)`` cell step
    (common/model.py:219-226) on the embedder's latent pair, as two 3xTF32 tensor-core GEMMs (``gi``, ``gh``: [N, 3D])
    between ``tpp_gru_mask_split`` and ``tpp_gru_gates`` (csrc/gru.cu).  There is no backward pass: the reference's
    ``optimize`` does not call the GRU (agents/ppo.py:116-121), its parameters never receive a gradient."""

    _tc = MLPEngineTC._tc

    def __init__(self, policy):
        assert policy.flat is not None and policy.recurrent
        self.policy, self.precision, self.n_launches = policy, 3, 0
        self.device = policy.flat.device
        self.D = D = policy.embedder.output_dim
        self.ld = _ceil(D, 32) * 32
        f = dict(dtype=torch.float32, device=self.device)
        self.off = {k: policy.layout[f"gru.gru.{k}"][0] for k in ("weight_ih_l0", "weight_hh_l0", "bias_ih_l0",
                                                                   "bias_hh_l0")}
        self.w = {k: (torch.zeros(3 * D, self.ld, **f), torch.zeros(3 * D, self.ld, **f)) for k in ("ih", "hh")}
        self._ws = {}
        self.refresh_weights()

    def _p(self, off):
        return _lib.C.c_void_p(self.policy.flat.data_ptr() + 4 * off)

    def refresh_weights(self):
        for k in ("ih", "hh"):
            _lib.call("tpp_split_tf32", self._p(self.off[f"weight_{k}_l0"]), self.D, 3 * self.D, self.D,
                      _lib.ptr(self.w[k][0]), _lib.ptr(self.w[k][1]), self.ld, None, None, 0, _lib.stream_ptr())
        self.n_launches += 2

    def step(self, x_pair, ldx, h_prev, done, h_out, N, slot=0):
        """x_pair: the latent ((hi, lo), row stride ldx); h_prev / h_out: [N, D] fp32 rows (may be the same tensor);
        done: uint8 [>= N] of the PREVIOUS env step (or None).  Returns ((hi, lo), ld) of h' for ``head_gemm``."""
        D, ld, s = self.D, self.ld, _lib.stream_ptr()
        ws = self._ws.get((N, slot))
        if ws is None:
            f = dict(dtype=torch.float32, device=self.device)
            ws = self._ws[(N, slot)] = dict(hm=(torch.zeros(N, ld, **f), torch.zeros(N, ld, **f)),
                                            out=(torch.zeros(N, ld, **f), torch.zeros(N, ld, **f)),
                                            gi=torch.zeros(N, 3 * D, **f), gh=torch.zeros(N, 3 * D, **f))
        assert h_prev.stride(-1) == 1 and h_out.stride(-1) == 1
        _lib.call("tpp_gru_mask_split", _lib.ptr(h_prev), h_prev.stride(0), _lib.ptr(done), N, D, _lib.ptr(ws["hm"][0]),
                  _lib.ptr(ws["hm"][1]), ld, s)
        bn = 64 if N * 3 * D <= 148 * 128 * 128 // 2 else 0
        self._tc(x_pair, ldx, self.w["ih"], ld, N, 3 * D, D, flags=EPI_BIAS, bias=self._p(self.off["bias_ih_l0"]),
                 out=ws["gi"], ldc=3 * D, block_n=bn)
        self._tc(ws["hm"], ld, self.w["hh"], ld, N, 3 * D, D, flags=EPI_BIAS, bias=self._p(self.off["bias_hh_l0"]),
                 out=ws["gh"], ldc=3 * D, block_n=bn)
        _lib.call("tpp_gru_gates", _lib.ptr(ws["gi"]), _lib.ptr(ws["gh"]), 3 * D, _lib.ptr(h_prev), h_prev.stride(0),
                  _lib.ptr(done), N, D, _lib.ptr(h_out), h_out.stride(0), _lib.ptr(ws["out"][0]), _lib.ptr(ws["out"][1]),
                  ld, s)
        self.n_launches += 2
        return ws["out"], ld
