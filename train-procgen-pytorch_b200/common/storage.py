"""GPU-resident rollout storage with the reference's ``Storage`` interface (common/storage.py:7-162).

Layout in HBM: per-step scalars are ``[T(+1), ld]`` rows (ld = N rounded up to 4);
vector observations are feature-major ``[T+1, n_obs, ld]`` so that the env kernels and the policy read/write
coalesced columns; image observations are uint8 NHWC ``[T+1, N, H, W, C]`` and become float NCHW/255 only when
a minibatch is gathered.  The reference-shaped public tensors (``obs_batch``, ``act_batch`` ...) are exposed as
views / on-demand conversions of these buffers.

GAE, advantage normalisation and the minibatch gather are CUDA kernels (csrc/storage.cu); minibatch indices come
from ``torch.randperm`` on the default CPU generator exactly like the reference's ``SubsetRandomSampler`` /
``BatchSampler(drop_last=True)`` pair, so minibatch composition is bit-identical under the same torch seed.
"""
from __future__ import annotations

from collections import deque

import numpy as np
import torch

from .. import _lib, parallel


def _round_up(x, m):
    return (x + m - 1) // m * m


class MiniBatch:
    """Device buffers of one gathered minibatch (reused between minibatches: no allocation in the loop)."""

    def __init__(self, mb, obs_width, ld_obs, device, split=False):
        """split: False = plain fp32 observations; True = TF32 pair (obs = hi, obs_lo = lo) for the tensor-core policy;
        "raw" (image observations only) = the integer pixel values 0..255 as fp32, exact in TF32, no lo half."""
        f = dict(dtype=torch.float32, device=device)
        self.mb, self.obs_width, self.ld_obs = mb, obs_width, ld_obs
        self.raw = split == "raw"
        self.obs = torch.zeros(mb, ld_obs, **f)
        self.obs_lo = torch.zeros(mb, ld_obs, **f) if split is True else None
        self.act = torch.zeros(mb, dtype=torch.int32, device=device)
        self.logp, self.value, self.ret, self.adv, self.done = (torch.zeros(mb, **f) for _ in range(5))


class Storage:
    def __init__(self, obs_shape, hidden_state_size, num_steps, num_envs, device, continuous_actions=False,
                 act_shape=None):
        if continuous_actions:
            raise NotImplementedError("continuous actions are outside the north-star hot path")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.TppError("Storage is GPU-resident: device must be a CUDA device (there is no CPU fallback)")
        _lib.load()
        self.continuous_actions = False
        self.performance_track = {}
        self.obs_shape = tuple(obs_shape)
        self.act_shape = act_shape
        self.hidden_state_size = hidden_state_size
        self.num_steps, self.num_envs = int(num_steps), int(num_envs)
        self.ld = _round_up(self.num_envs, 32)   # row stride of every [T][ld] buffer (TMA / MN-major operand friendly)
        self.is_image = len(self.obs_shape) == 3
        self.obs_width = int(np.prod(self.obs_shape))
        self.world_size, self.process_group = 1, None
        # "exact": sequential recurrence in the reference's fp32 operation order (raw advantages / returns bit-identical
        # to torch CPU); "warp_scan": warp-level segmented scan over n_steps fused with moments + normalisation in one
        # launch (tpp_gae_scan; agrees to ~4e-7 of the advantage scale, stated tolerance 1e-5)
        self.gae_mode = "exact"
        self._mb = {}
        self._pinned = None
        self.h2d_bytes = 0       # bytes staged from the host by store()/stage_step() (bench accounting)
        self.n_launches = 0
        self.reset()

    # ------------------------------------------------------------------------------------------
    def reset(self):
        T, N, ld, dev = self.num_steps, self.num_envs, self.ld, self.device
        f = dict(dtype=torch.float32, device=dev)
        if self.is_image:
            c, h, w = self.obs_shape
            self.frames = torch.zeros(T + 1, N, h, w, c, dtype=torch.uint8, device=dev)
            self.obs_fm = None
        else:
            self.obs_fm = torch.zeros(T + 1, self.obs_width, ld, **f)
            self.frames = None
        self.act_i32 = torch.zeros(T, ld, dtype=torch.int32, device=dev)
        self.logp = torch.zeros(T, ld, **f)
        self.rew = torch.zeros(T, ld, **f)
        self.env_rew = None            # raw rewards when the env normalises them (Box-World / Procgen)
        self.done_u8 = torch.zeros(T, ld, dtype=torch.uint8, device=dev)
        self.value = torch.zeros(T + 1, ld, **f)
        self.ret = torch.zeros(T, ld, **f)
        self.adv = torch.zeros(T, ld, **f)
        self.moments4 = torch.zeros(4, dtype=torch.float64, device=dev)   # [3]: grid-barrier counter of tpp_gae_scan
        self.moments = self.moments4[:3]
        self.info_batch = deque(maxlen=T)
        self._hidden = None
        self.done_carry = torch.zeros(ld, dtype=torch.uint8, device=dev)   # done flags of the previous rollout's last step
        self.step = 0

    # ---- reference-shaped public tensors -------------------------------------------------------------
    @property
    def obs_batch(self):
        if self.is_image:
            return self.frames.permute(0, 1, 4, 2, 3).float() / 255.0
        return self.obs_fm.permute(0, 2, 1)[:, :self.num_envs]

    @property
    def hidden_states_batch(self):
        if self._hidden is None:
            self._hidden = torch.zeros(self.num_steps + 1, self.num_envs, self.hidden_state_size, device=self.device)
        return self._hidden

    @property
    def act_batch(self):
        return self.act_i32[:, :self.num_envs].float()

    @property
    def log_prob_act_batch(self):
        return self.logp[:, :self.num_envs]

    @property
    def rew_batch(self):
        return self.rew[:, :self.num_envs]

    @property
    def done_batch(self):
        return self.done_u8[:, :self.num_envs].float()

    @property
    def value_batch(self):
        return self.value[:, :self.num_envs]

    @property
    def return_batch(self):
        return self.ret[:, :self.num_envs]

    @property
    def adv_batch(self):
        return self.adv[:, :self.num_envs]

    # ---- rollout slots for the fused kernels -----------------------------------------------------------
    def obs_slot(self, t):
        return self.frames[t] if self.is_image else self.obs_fm[t]

    def enable_raw_rewards(self):
        if self.env_rew is None:
            self.env_rew = torch.zeros(self.num_steps, self.ld, dtype=torch.float32, device=self.device)
            # integer rewards as the env kernel writes them (normalised once per rollout, tpp_vecnormalize_rollout)
            self.env_rew_i32 = torch.zeros(self.num_steps, self.ld, dtype=torch.int32, device=self.device)

    # ---- reference API: host-driven stores (compat path) -------------------------------------------------
    def _t(self, x, dtype):
        if torch.is_tensor(x):
            return x.to(self.device, dtype)
        return torch.from_numpy(np.ascontiguousarray(x)).to(self.device, dtype)

    def _store_obs(self, slot, obs):
        if self.is_image and getattr(obs, "dtype", None) in (np.uint8, torch.uint8):
            # raw uint8 NHWC frames (what the Procgen engine emits before Transpose/Scale): staged through DOUBLE-BUFFERED
            # pinned memory straight into the rollout slot, 4x less PCIe traffic than the float NCHW contract.  The host
            # fills buffer k % 2 while the copy out of buffer (k - 1) % 2 may still be in flight; a buffer is reused only
            # after the event behind its last copy has completed (an event wait, not a stream synchronisation).
            src = obs if torch.is_tensor(obs) else torch.from_numpy(np.ascontiguousarray(obs))
            assert tuple(src.shape) == tuple(self.frames[slot].shape), (src.shape, self.frames[slot].shape)
            if not src.is_cuda:
                if self._pinned is None:
                    self._pinned = [[torch.empty_like(src).pin_memory(), None] for _ in range(2)]
                    self._pinned_k = 0
                buf = self._pinned[self._pinned_k]
                self._pinned_k ^= 1
                if buf[1] is not None:
                    buf[1].synchronize()
                buf[0].copy_(src)
                self.frames[slot].copy_(buf[0], non_blocking=True)
                buf[1] = torch.cuda.Event()
                buf[1].record()
            else:
                self.frames[slot].copy_(src, non_blocking=True)
            self.h2d_bytes += src.numel()
            return
        o = self._t(obs, torch.float32)
        self.h2d_bytes += 0 if (torch.is_tensor(obs) and obs.is_cuda) else o.numel() * 4
        if self.is_image:
            self.frames[slot] = (o * 255.0).round().clamp(0, 255).to(torch.uint8).permute(0, 2, 3, 1)
        else:
            self.obs_fm[slot, :, :self.num_envs] = o.reshape(self.num_envs, -1).t()

    def store(self, obs, hidden_state, act, rew, done, info, log_prob_act, value):
        s, N = self.step, self.num_envs
        self._store_obs(s, obs)
        if self.hidden_state_size and hidden_state is not None and np.size(hidden_state):
            self.hidden_states_batch[s] = self._t(hidden_state, torch.float32).reshape(N, -1)
        self.act_i32[s, :N] = self._t(act, torch.int32).reshape(-1)
        self.rew[s, :N] = self._t(rew, torch.float32).reshape(-1)
        self.done_u8[s, :N] = self._t(done, torch.uint8).reshape(-1)
        self.logp[s, :N] = self._t(log_prob_act, torch.float32).reshape(-1)
        self.value[s, :N] = self._t(value, torch.float32).reshape(-1)
        self.info_batch.append(info)
        self.step = (self.step + 1) % self.num_steps

    def stage_obs(self, slot, obs):
        """Host-env staging: copy one observation batch (uint8 NHWC frames, float NCHW, or [N, n_obs]) into a slot."""
        self._store_obs(slot, obs)

    def _host_rows(self):
        if getattr(self, "_host_step", None) is None:
            T, ld = self.num_steps, self.ld
            self._host_step = dict(rew=torch.zeros(T, ld, dtype=torch.float32).pin_memory(),
                                   raw=torch.zeros(T, ld, dtype=torch.float32).pin_memory(),
                                   done=torch.zeros(T, ld, dtype=torch.uint8).pin_memory(),
                                   act=torch.zeros(self.num_envs, dtype=torch.int32).pin_memory(), has_raw=False)
        return self._host_step

    def start_action_fetch(self, slot):
        """Host-env staging: start the device -> host copy of the N actions the policy drew into slot ``slot`` (pinned,
        asynchronous) and record ITS event."""
        h = self._host_rows()
        h["act"].copy_(self.act_i32[slot, :self.num_envs], non_blocking=True)
        h["act_event"] = torch.cuda.Event()
        h["act_event"].record()

    def finish_action_fetch(self):
        """Wait for that copy alone (an event wait: the host never synchronises the whole stream) -> int64 numpy."""
        h = self._host_rows()
        h["act_event"].synchronize()
        self.d2h_bytes = getattr(self, "d2h_bytes", 0) + 4 * self.num_envs
        return h["act"].numpy().astype(np.int64)

    def fetch_actions(self, slot):
        self.start_action_fetch(slot)
        return self.finish_action_fetch()

    def stage_step(self, slot, rew, done, info=None, raw_rew=None, upload_done=False):
        """Host-env staging of one step's reward / done vectors (the action, log-prob and value are already in the rollout:
        they were produced on the device).  Rows collect in pinned host memory and go up in ONE copy per rollout
        (``flush_steps``).  The raw env reward the logger wants (common/storage.py:131-137) is ``raw_rew`` or, like the
        reference, ``info[i]['env_reward']`` when the env's VecNormalize wrapper put it there; ``info[i]['env_done']``
        likewise overrides ``done`` for the logger's episode accounting only when present."""
        h, N = self._host_rows(), self.num_envs
        h["rew"][slot, :N] = torch.as_tensor(np.asarray(rew, dtype=np.float32).reshape(-1))
        h["done"][slot, :N] = torch.as_tensor(np.asarray(done).reshape(-1).astype(np.uint8))
        if raw_rew is None and info is not None and len(info) == N and isinstance(info[0], dict) \
                and "env_reward" in info[0]:
            raw_rew = np.fromiter((i["env_reward"] for i in info), dtype=np.float32, count=N)
        if raw_rew is not None:
            h["raw"][slot, :N] = torch.as_tensor(np.asarray(raw_rew, dtype=np.float32).reshape(-1))
            h["has_raw"] = True
        if upload_done:      # recurrent policies mask the hidden state of step slot + 1 with these flags: N bytes now
            self.done_u8[slot].copy_(h["done"][slot], non_blocking=True)
            self.h2d_bytes += N

    def flush_steps(self):
        """One H2D copy per rollout of the staged reward / done (/ raw reward) rows."""
        h = self._host_rows()
        self.rew.copy_(h["rew"], non_blocking=True)
        self.done_u8.copy_(h["done"], non_blocking=True)
        self.h2d_bytes += self.num_steps * self.num_envs * 5
        if h["has_raw"]:
            if self.env_rew is None:
                self.enable_raw_rewards()
            self.env_rew.copy_(h["raw"], non_blocking=True)
            self.h2d_bytes += self.num_steps * self.num_envs * 4

    def store_last(self, last_obs, last_hidden_state, last_value):
        self._store_obs(self.num_steps, last_obs)
        if self.hidden_state_size and last_hidden_state is not None and np.size(last_hidden_state):
            self.hidden_states_batch[self.num_steps] = self._t(last_hidden_state, torch.float32).reshape(self.num_envs, -1)
        self.value[self.num_steps, :self.num_envs] = self._t(last_value, torch.float32).reshape(-1)

    # ---- GAE -----------------------------------------------------------------------------------------------
    def compute_estimates(self, gamma=0.99, lmbda=0.95, use_gae=True, normalize_adv=True):
        if not use_gae:
            # common/storage.py:69-77: the reference overwrites the Monte-Carlo returns with adv(=0)+V, i.e. the
            # use_gae=False branch is broken upstream; refuse rather than silently reproduce or "fix" it.
            raise NotImplementedError("use_gae=False is broken in the reference (storage.py:69-77); unsupported")
        T, N, s = self.num_steps, self.num_envs, _lib.stream_ptr()
        self.moments4.zero_()
        if self.gae_mode == "warp_scan":
            fuse = 1 if (normalize_adv and self.world_size == 1) else 0
            if _lib.try_call("tpp_gae_scan", _lib.ptr(self.rew), _lib.ptr(self.done_u8), _lib.ptr(self.value),
                             _lib.ptr(self.adv), _lib.ptr(self.ret), _lib.ptr(self.moments4), T, N, self.ld, float(gamma),
                             float(lmbda), fuse, s):
                self.n_launches += 1
                if normalize_adv and not fuse:
                    parallel.allreduce_moments_(self.moments, self.process_group)
                    _lib.call("tpp_adv_normalize", _lib.ptr(self.adv), _lib.ptr(self.moments), T, N, self.ld, s)
                    self.n_launches += 1
                return
            # (ENOTSUP: T or N outside the fused kernel's range -> the exact kernels below)
        elif self.gae_mode != "exact":
            raise ValueError(f"gae_mode must be 'exact' or 'warp_scan', not {self.gae_mode!r}")
        _lib.call("tpp_gae", _lib.ptr(self.rew), _lib.ptr(self.done_u8), _lib.ptr(self.value), _lib.ptr(self.adv),
                  _lib.ptr(self.ret), _lib.ptr(self.moments), T, N, self.ld, float(gamma), float(lmbda), s)
        self.n_launches += 1
        if normalize_adv:
            if self.world_size > 1:   # exact global moments under env sharding: 3 doubles, once per rollout
                parallel.allreduce_moments_(self.moments, self.process_group)
            _lib.call("tpp_adv_normalize", _lib.ptr(self.adv), _lib.ptr(self.moments), T, N, self.ld, s)
            self.n_launches += 1

    # ---- minibatches -----------------------------------------------------------------------------------------
    def minibatch_buffers(self, mb, ld_obs=None, split=False, slot=0):
        ld_obs = ld_obs or _round_up(self.obs_width, 4)
        key = (mb, ld_obs, split, slot)
        if key not in self._mb:
            self._mb[key] = MiniBatch(mb, self.obs_width, ld_obs, self.device, split)
        return self._mb[key]

    @staticmethod
    def randperm(n, out=None):
        """``torch.randperm(n)`` on the default CPU generator -- the same permutation AND the same generator state
        afterwards -- through the library's host restatement (``tpp_randperm_mt19937``: MT19937 + ATen's Fisher-Yates
        loop on a 4-byte index array, ~2x faster; three million-element draws per iteration were the host-side
        bottleneck of ``PPO.train``).  Falls back to torch when the state blob is not the layout this build knows."""
        state = torch.get_rng_state()
        if state.numel() != 5056 or n < 2 or n >= (1 << 32) // 20:
            perm = torch.randperm(n)
            if out is not None:
                out.copy_(perm)
                return out
            return perm
        a = state.numpy()           # [0:8) seed, [8:12) left, [12:16) seeded, [16:24) next, [24:5016) 624 words (u64 each)
        out = torch.empty(n, dtype=torch.int64) if out is None else out
        assert out.dtype in (torch.int64, torch.int32) and out.is_contiguous() and out.numel() == n and not out.is_cuda
        _lib.call("tpp_randperm_mt19937" if out.dtype == torch.int64 else "tpp_randperm_mt19937_i32",
                  a[24:24 + 624 * 8].ctypes.data, a[8:12].ctypes.data, a[16:24].ctypes.data, int(n), out.data_ptr())
        torch.set_rng_state(state)
        return out

    def epoch_indices(self, mini_batch_size):
        """One ``torch.randperm(T*N)`` on the default CPU generator, cut into consecutive minibatches
        (drop_last) — common/storage.py:87-91.  Returns a device int64 tensor [n_mb, mb]."""
        batch = self.num_steps * self.num_envs
        n_mb = batch // mini_batch_size
        perm = self.randperm(batch)
        self.last_perm = perm
        idx = perm[:n_mb * mini_batch_size].view(n_mb, mini_batch_size)
        return idx.pin_memory().to(self.device, non_blocking=True)

    @staticmethod
    def recurrent_perm(num_steps, num_envs, mini_batch_size):
        """Minibatches of a recurrent policy (common/storage.py:93-110): ONE ``torch.randperm(num_envs)`` on the default
        CPU generator, ``num_envs // (T*N // mini_batch_size)`` whole trajectories per minibatch, rows flattened
        time-major like ``batch[:, idxes].reshape(-1)``.  Host side: returns (int64 [n_batches, T * envs_per_batch] flat
        indices t*N + e, int64 [n_batches, envs_per_batch] env indices).  Ragged last batches (num_envs not a multiple
        of the per-batch env count) are refused: every kernel of the update is shaped by the minibatch size."""
        per_epoch = (num_steps * num_envs) // mini_batch_size
        if per_epoch < 1 or num_envs // per_epoch < 1:
            raise ValueError("recurrent minibatches need mini_batch_size <= T*N and at least one env per minibatch")
        envs_per_batch = num_envs // per_epoch
        if num_envs % envs_per_batch:
            raise NotImplementedError(f"ragged recurrent minibatches ({num_envs} envs in batches of {envs_per_batch})")
        envs = torch.randperm(num_envs).view(-1, envs_per_batch)
        flat = (torch.arange(num_steps).view(1, -1, 1) * num_envs + envs.view(-1, 1, envs_per_batch))
        return flat.reshape(envs.shape[0], -1), envs

    def epoch_indices_recurrent(self, mini_batch_size):
        """``recurrent_perm`` for this storage, on the device ([n_batches, rows] int64): rows of ``gather``."""
        flat, envs = self.recurrent_perm(self.num_steps, self.num_envs, mini_batch_size)
        self.last_env_perm = envs
        return flat.pin_memory().to(self.device, non_blocking=True)

    def epoch_perm_pinned(self, mini_batch_size, slot):
        """The same draw as ``epoch_indices`` left in a reusable pinned host buffer (``slot``): the caller uploads it
        on a copy stream while the previous epoch's kernels run.  Returns (pinned int32 [n_mb, mb], ready event or
        None): the event of the last upload from this buffer must have completed before it is overwritten.  The buffer is
        int32 (the device widens it to the gather's int64 index buffer in the copy it makes anyway)."""
        batch = self.num_steps * self.num_envs
        n_mb = batch // mini_batch_size
        pool = self.__dict__.setdefault("_perm_pool", {})
        key = (slot, n_mb, mini_batch_size)
        if key not in pool:         # int32: T*N < 2^31 for every config; half the pinned bytes and H2D traffic
            pool[key] = [torch.empty(n_mb, mini_batch_size, dtype=torch.int32).pin_memory(), None]
        buf, busy = pool[key]
        if busy is not None:
            busy.synchronize()                      # the previous upload from this buffer has left the host
        if n_mb * mini_batch_size == batch:
            perm = self.randperm(batch, out=buf.view(-1))          # drawn straight into the pinned buffer
        else:
            perm = self.randperm(batch)
            buf.view(-1).copy_(perm[:n_mb * mini_batch_size])
        self.last_perm = perm
        return buf, pool[key]

    def gather(self, idx_row, out):
        """Gather one minibatch (device int64 indices [mb]) into ``out`` (a MiniBatch)."""
        s, N = _lib.stream_ptr(), self.num_envs
        scal = (_lib.ptr(self.act_i32), _lib.ptr(self.logp), _lib.ptr(self.value), _lib.ptr(self.ret),
                _lib.ptr(self.adv), _lib.ptr(self.done_u8))
        outs = (_lib.ptr(out.act), _lib.ptr(out.logp), _lib.ptr(out.value), _lib.ptr(out.ret), _lib.ptr(out.adv),
                _lib.ptr(out.done))
        # flat index k = t*N + e is decoded in-kernel; scalar rows are ld apart
        if self.is_image:
            c, h, w = self.obs_shape
            _lib.call("tpp_gather_img", _lib.ptr(idx_row), out.mb, N, self.ld, h, w, c, _lib.ptr(self.frames), *scal,
                      _lib.ptr(out.obs), _lib.ptr(out.obs_lo), out.ld_obs, *outs, 1 if out.raw else 0, s)
        else:
            assert not out.raw, "raw pixel mode is for image observations"
            _lib.call("tpp_gather_vec", _lib.ptr(idx_row), out.mb, N, self.ld, self.obs_width, _lib.ptr(self.obs_fm),
                      *scal, _lib.ptr(out.obs), _lib.ptr(out.obs_lo), out.ld_obs, *outs, s)
        self.n_launches += 1
        return out

    def fetch_train_generator(self, mini_batch_size=None, recurrent=False):
        batch = self.num_steps * self.num_envs
        mini_batch_size = mini_batch_size or batch
        if recurrent:       # whole trajectories of permuted envs; only the INITIAL hidden states travel (storage.py:99-102)
            idx = self.epoch_indices_recurrent(mini_batch_size)
            for i in range(idx.shape[0]):
                rows = idx.shape[1]
                out = self.gather(idx[i], MiniBatch(rows, self.obs_width, _round_up(self.obs_width, 4), self.device))
                obs = out.obs[:, :self.obs_width].reshape(rows, *self.obs_shape)
                hidden = self.hidden_states_batch[0, self.last_env_perm[i].to(self.device)]
                yield obs, hidden, out.act.float(), out.done, out.logp, out.value, out.ret, out.adv
            return
        idx = self.epoch_indices(mini_batch_size)
        hidden = torch.zeros(1, self.hidden_state_size, device=self.device).expand(batch, self.hidden_state_size)
        for i in range(idx.shape[0]):
            out = self.gather(idx[i], MiniBatch(mini_batch_size, self.obs_width, _round_up(self.obs_width, 4),
                                                self.device))
            obs = out.obs[:, :self.obs_width].reshape(mini_batch_size, *self.obs_shape)
            yield obs, hidden, out.act.float(), out.done, out.logp, out.value, out.ret, out.adv

    # ---- logging ---------------------------------------------------------------------------------------------
    def snapshot_log_data(self):
        """Start the device -> host copy of this rollout's (raw reward, done) batches into pinned buffers on the current
        stream and return a closure that waits for it and yields ``fetch_log_data()``'s triple.  Lets ``PPO.train``
        launch the next rollout before the host touches the log data (the rollout overwrites these buffers)."""
        rew = self.env_rew if self.env_rew is not None else self.rew
        if getattr(self, "_log_host", None) is None:
            self._log_host = (torch.empty(rew.shape, dtype=torch.float32).pin_memory(),
                              torch.empty(self.done_u8.shape, dtype=torch.uint8).pin_memory())
        self._log_host[0].copy_(rew, non_blocking=True)
        self._log_host[1].copy_(self.done_u8, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        N = self.num_envs

        def finish():
            ev.synchronize()
            return (self._log_host[0][:, :N].numpy().copy(), self._log_host[1][:, :N].numpy().astype(np.float32),
                    float("nan"))
        return finish

    def snapshot_episodes(self, keep=40):
        """Device-side episode accounting (replaces fetch_log_data + the loop of Logger.feed,
        common/storage.py:130-162, common/logger.py:119-147): ``tpp_episode_scan`` closes the episodes of this rollout on
        the device -- open-episode returns / lengths persist in HBM between rollouts -- and only the episode count and
        the last ``keep`` (return, length) records in env-major order start their way to a pinned host buffer.  Returns a
        closure that waits for the copy and yields that float64 record (``Logger.feed_episodes``)."""
        rew = self.env_rew if self.env_rew is not None else self.rew
        N, dev = self.num_envs, self.device
        if getattr(self, "_ep_state", None) is None or self._ep_state[3].numel() != 2 + 2 * keep:
            self._ep_state = (torch.zeros(N, dtype=torch.float64, device=dev),
                              torch.zeros(N, dtype=torch.int32, device=dev),
                              torch.zeros(N + 1, dtype=torch.int32, device=dev),
                              torch.zeros(2 + 2 * keep, dtype=torch.float64, device=dev))
            self._ep_host = torch.zeros(2 + 2 * keep, dtype=torch.float64).pin_memory()
        run_ret, run_len, scratch, out = self._ep_state
        _lib.call("tpp_episode_scan", _lib.ptr(rew), _lib.ptr(self.done_u8), self.num_steps, N, self.ld,
                  _lib.ptr(run_ret), _lib.ptr(run_len), _lib.ptr(scratch), _lib.ptr(out), keep, _lib.stream_ptr())
        self.n_launches += 3
        self._ep_host.copy_(out, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()

        def finish():
            ev.synchronize()
            return self._ep_host.numpy().copy()
        return finish

    def fetch_log_data(self):
        """(rew_batch [T,N], done_batch [T,N], true_average_reward) as numpy, raw env rewards when available
        (common/storage.py:130-162; per-level tracking needs Procgen's prev_level_seed and stays NaN here)."""
        rew = self.env_rew if self.env_rew is not None else self.rew
        return (rew[:, :self.num_envs].cpu().numpy(), self.done_u8[:, :self.num_envs].float().cpu().numpy(),
                float("nan"))
