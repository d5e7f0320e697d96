"""PPO agent with the reference's interface (agents/ppo.py:9-279) on B200-native kernels.

``optimize()`` = epochs x minibatches of: device gather -> policy forward -> fused loss fwd+bwd -> policy
backward into the flat gradient buffer -> (one NCCL all-reduce when sharded) -> fused clip+Adam.  No autograd,
no host synchronisation inside the loop: per-minibatch loss statistics are accumulated on the device and read
back once per call.  ``train()`` runs the rollout on the device (policy forward -> Philox action sampling ->
fused env step writing into the rollout slots), optionally replayed from a CUDA graph.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from .. import _lib, parallel
from ..common.engine import GRUCellTC, ImpalaEngineTC, MLPEngine, MLPEngineTC
from ..common.model import ImpalaModel, MLPModel
from .base_agent import BaseAgent


def adjust_lr(optimizer, init_lr, timesteps, max_timesteps):
    """common/misc_util.py:92-96."""
    lr = init_lr * (1 - (timesteps / max_timesteps))
    for g in optimizer.param_groups:
        g["lr"] = lr
    return optimizer, lr


def adjust_lr_grok(optimizer, init_lr, timesteps, max_timesteps):
    """common/misc_util.py:98-102."""
    lr = init_lr * (1.1 ** (timesteps / 1e6))
    for g in optimizer.param_groups:
        g["lr"] = lr
    return optimizer, lr


class FlatAdam:
    """Adam(lr, betas=(0.9, 0.999), eps) + clip_grad_norm_ over the policy's flat buffer; state lives on the
    device (``tpp_adam_state``).  ``state_dict()`` has torch.optim.Adam's layout so reference tooling can load it."""

    def __init__(self, policy, lr, eps=1e-5, betas=(0.9, 0.999), max_grad_norm=0.5):
        self.policy = policy
        self.flat, self.gflat = policy.flat, policy.flat_grad
        self.m = torch.zeros_like(self.flat)
        self.v = torch.zeros_like(self.flat)
        self.param_groups = [dict(lr=lr, betas=betas, eps=eps, weight_decay=0, amsgrad=False)]
        host = _lib.AdamState(lr=lr, beta1=betas[0], beta2=betas[1], eps=eps, max_grad_norm=max_grad_norm,
                              grad_scale=1.0, step=0)
        raw = bytes(host)
        self.state = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(self.flat.device)
        self._f64 = self.state.view(torch.float64)    # [lr, beta1, beta2, eps, (f32 pair), (i32 pair), sqnorm x2]
        self._f32 = self.state.view(torch.float32)    # max_grad_norm @8, grad_scale @9
        self._i32 = self.state.view(torch.int32)      # step @10, ticket @11
        self._lr_on_device = lr
        self.n_launches = 0
        self.peer = None          # parallel.PeerReducer under env sharding

    def set_grad_scale(self, s):
        self._f32[9:10].copy_(torch.tensor([s], dtype=torch.float32))

    def _sync_lr(self):
        lr = float(self.param_groups[0]["lr"])
        if lr != self._lr_on_device:
            # pinned staging: an async copy from PAGEABLE memory first synchronises the stream, i.e. the host would wait
            # for the whole rollout here and launch the epoch graph into an idle GPU
            if getattr(self, "_lr_pinned", None) is None:
                self._lr_pinned = torch.zeros(1, dtype=torch.float64).pin_memory()
            self._lr_pinned[0] = lr
            self._f64[0:1].copy_(self._lr_pinned, non_blocking=True)
            self._lr_on_device = lr

    def step(self, views=None):
        self._sync_lr()
        self.launch(views)

    def launch(self, views=None):
        """The two kernels of a step (capturable in a CUDA graph; ``_sync_lr`` must have run outside the capture).  With a
        peer reducer (env sharding on one node) the first one is the one-shot all-reduce over NVLink peer memory fused
        with the norm reduction, and Adam runs on the reduced copy."""
        n, s = self.flat.numel(), _lib.stream_ptr()
        g = self.gflat
        if self.peer is not None:
            self.peer.launch(self.gflat, self.state)
            g = self.peer.reduced
        else:
            _lib.call("tpp_grad_sqnorm", _lib.ptr(self.state), _lib.ptr(self.gflat), n, s)
        # ``views``: an engine's (tpp_weight_view array, count): the step also rewrites the tensor-core operand copies
        arr, cnt = views if views is not None else (None, 0)
        _lib.call("tpp_adam_clip_step_views", _lib.ptr(self.state), _lib.ptr(self.flat), _lib.ptr(g),
                  _lib.ptr(self.m), _lib.ptr(self.v), n, arr, cnt, s)
        self.n_launches += 2

    @property
    def step_count(self):
        return int(self._i32[10].item())

    def state_dict(self):
        state, names = {}, [n for n, _ in self.policy.named_parameters()]
        step = torch.tensor(float(self.step_count))
        for i, n in enumerate(names):
            off, shape = self.policy.layout[n]
            k = int(np.prod(shape))
            state[i] = {"step": step.clone(), "exp_avg": self.m[off:off + k].view(shape).clone(),
                        "exp_avg_sq": self.v[off:off + k].view(shape).clone()}
        group = dict(self.param_groups[0], params=list(range(len(names))))
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd):
        names = [n for n, _ in self.policy.named_parameters()]
        step = 0
        for i, n in enumerate(names):
            if i in sd["state"]:
                off, shape = self.policy.layout[n]
                k = int(np.prod(shape))
                self.m[off:off + k].copy_(sd["state"][i]["exp_avg"].reshape(-1))
                self.v[off:off + k].copy_(sd["state"][i]["exp_avg_sq"].reshape(-1))
                step = int(sd["state"][i]["step"])
        self._i32[10:11].copy_(torch.tensor([step], dtype=torch.int32))
        self.param_groups[0]["lr"] = sd["param_groups"][0]["lr"]


class PPO(BaseAgent):
    def __init__(self, env, policy, logger, storage, device, n_checkpoints, env_valid=None, storage_valid=None,
                 n_steps=128, n_envs=8, epoch=3, n_minibatch=8, mini_batch_size=32 * 8, gamma=0.99, lmbda=0.95,
                 learning_rate=2.5e-4, grad_clip_norm=0.5, eps_clip=0.2, value_coef=0.5, entropy_coef=0.01,
                 x_entropy_coef=0., normalize_adv=True, normalize_rew=True, use_gae=True, entropy_scaling=None,
                 increasing_lr=False, sparsity_coef=0., fs_coef=0., **kwargs):
        super().__init__(env, policy, logger, storage, device, n_checkpoints, env_valid, storage_valid)
        self.fs_coef = fs_coef
        self.total_timesteps = 0
        self.entropy_scaling = entropy_scaling
        self.entropy_multiplier = 1.
        self.s_loss_coef = sparsity_coef
        self.min_rew, self.max_rew = -1., 11.
        self.n_steps, self.n_envs = n_steps, n_envs
        self.epoch, self.n_minibatch, self.mini_batch_size = epoch, n_minibatch, mini_batch_size
        self.gamma, self.lmbda = gamma, lmbda
        self.learning_rate = learning_rate
        self.grad_clip_norm, self.eps_clip = grad_clip_norm, eps_clip
        self.value_coef, self.entropy_coef, self.x_entropy_coef = value_coef, entropy_coef, x_entropy_coef
        self.normalize_adv, self.normalize_rew, self.use_gae = normalize_adv, normalize_rew, use_gae
        self.adjust_lr = adjust_lr_grok if increasing_lr else adjust_lr
        self.use_cuda_graph = bool(kwargs.get("use_cuda_graph", True))
        self.sample_seed = int(kwargs.get("sample_seed", 0))
        self.use_epoch_graph = bool(kwargs.get("epoch_graph", True))    # False: per-group graphs (the multi-GPU path)
        # minibatches of one gradient-accumulation window processed as one pass ("auto" = all, int = cap, 1 = off)
        self.fuse_accum = kwargs.get("fuse_accum", "auto")
        # env ranges stepping through the rollout as concurrent kernel chains (an int or "auto"; default 1 = off:
        # measured on B200 at 4096 envs, 1 / 2 / 4 chains take the same time -- the rollout is T sequential steps of
        # dependent launches and a smaller range does not shorten a step; kept for larger env counts)
        self.rollout_chains = kwargs.get("rollout_chains", 1)
        # rollout: last embedder layer + heads + action draw in one CUDA-core launch (MLP policies, TC engine)
        self.fused_tail = bool(kwargs.get("fused_tail", True))
        # rollout: the WHOLE policy step (4 dense layers + heads + action draw) in one cluster kernel
        # (tpp_policy_rollout_fused); False keeps the per-layer GEMM launches + tail kernel
        self.fused_rollout = bool(kwargs.get("fused_rollout", True))
        self.max_group_rows = int(kwargs.get("max_group_rows", 1 << 18))
        # GAE: "exact" (bit-identical to the reference's sequential fp32 recurrence) or "warp_scan" (one fused launch:
        # warp-level segmented scan over n_steps + moments + normalisation, common/storage.py::compute_estimates)
        if "gae_mode" in kwargs:
            for st in (storage, storage_valid):
                if st is not None:
                    st.gae_mode = kwargs["gae_mode"]
        # sharded runs keep the whole-epoch graph: the per-step ncclAllReduce is captured with the kernels around it
        # (False: per-group graphs with the all-reduce launched from the host between them)
        self.graph_allreduce = bool(kwargs.get("graph_allreduce", True))
        self.peer_reduce = bool(kwargs.get("peer_reduce", True))

        if policy.flat is None:
            policy.flatten_(device)
        self.n_actions = policy.action_size
        # dense-layer arithmetic: "tf32x3" (default) = tcgen05 tensor cores with the error-compensated 3xTF32 split
        # (fp32-grade, the parity path), "tf32" = single-pass tensor-core fast mode (~1e-3), "fp32" = exact CUDA-core
        # GEMM (MLP only).  All three are this library's kernels; there is no cuDNN / cuBLAS route in the product
        # (tests inject their torch-autograd cross-check engine through ``engine=``).
        self.matmul = kwargs.get("matmul", "tf32x3")
        if self.matmul not in ("tf32x3", "tf32", "fp32"):
            raise ValueError(f"matmul must be tf32x3, tf32 or fp32, not {self.matmul!r}")
        if kwargs.get("engine") is not None:
            self.engine = kwargs["engine"]
        elif isinstance(policy.embedder, MLPModel):
            if self.matmul == "fp32":
                self.engine = MLPEngine(policy, self.n_actions)
            else:
                # image observations reach the first layer as integer pixel values (exact in TF32: no lo half, two
                # MMA passes) with ScaledFloatFrame's 1/255 folded into the layer's weight copy
                self.engine = MLPEngineTC(policy, self.n_actions, precision=3 if self.matmul == "tf32x3" else 1,
                                          raw_pixels=bool(getattr(storage, "is_image", False)),
                                          obs_shape=getattr(storage, "obs_shape", None))
        elif isinstance(policy.embedder, ImpalaModel):
            # IMPALA convolutions as implicit GEMMs on the tcgen05 kernel (TMA im2col)
            self.engine = ImpalaEngineTC(policy, self.n_actions, storage.obs_shape,
                                         precision=1 if self.matmul == "tf32" else 3)
        else:
            raise NotImplementedError(f"no engine for embedder {type(policy.embedder).__name__}: the hot path covers "
                                      "MLPModel and ImpalaModel (SURVEY 8a10)")
        # Recurrent policies (row N4): the GRU acts at prediction time only -- the reference's optimize() evaluates
        # embedder + heads without it (agents/ppo.py:116-121) on env-permuting minibatches (common/storage.py:93-110).
        self.recurrent = bool(policy.is_recurrent())
        self.gru = None
        if self.recurrent:
            if not isinstance(self.engine, (MLPEngineTC, ImpalaEngineTC)):
                raise NotImplementedError("recurrent policies run on the tensor-core engines (matmul='tf32x3' or 'tf32')")
            self.gru = GRUCellTC(policy)
            self.fused_rollout = self.fused_tail = False       # the cell sits between the embedder and the heads
            self.rollout_chains = 1
        self.optimizer = FlatAdam(policy, learning_rate, eps=1e-5, max_grad_norm=grad_clip_norm)
        self.world_size = 1
        self.process_group = None
        if storage_valid is not None:        # decorrelate the validation rollout's action draws from the training one's
            storage_valid.sample_offset = int(n_envs)
        self._tick = torch.zeros(1, dtype=torch.int64, device=policy.flat.device)
        self._rollout_graph = None
        self._stats = None
        self._pbar = None
        self.last_stats = None
        self.n_launches = 0

    # ------------------------------------------------------------------------------------------
    def shard(self, world_size, process_group=None, rank=None):
        """Env-sharded data parallel: gradients are summed with ONE all-reduce per optimizer step and scaled by
        1/world inside the clip+Adam kernel; advantage moments are all-reduced once per rollout.  The rank is folded
        into the action-sampling key (global env index = rank * n_envs + e), so shards never draw the same uniforms;
        env seeds are the caller's (use seed + rank)."""
        if rank is None:
            rank = torch.distributed.get_rank(process_group) if torch.distributed.is_initialized() else 0
        if getattr(self, "_shard_rank", 0) != rank:
            assert not self.__dict__.get("_graphs") and not self.__dict__.get("_host_graphs"), \
                "shard() must precede the first rollout (the sampling offset is baked into captured graphs)"
        self._shard_rank = rank
        for st in (self.storage, self.storage_valid):
            if st is not None:
                st.sample_offset = getattr(st, "sample_offset", 0) % (2 * self.n_envs) + 2 * self.n_envs * rank
        self.world_size, self.process_group = world_size, process_group
        self.optimizer.set_grad_scale(1.0 / world_size)
        # the step's only collective: one-shot all-reduce over NVLink peer memory fused with the norm reduction
        # (``peer_reduce=False`` keeps ncclAllReduce, captured in the epoch graph)
        if world_size > 1 and self.peer_reduce and torch.distributed.is_initialized() \
                and torch.distributed.get_backend(process_group) == "nccl":
            try:
                self.optimizer.peer = parallel.PeerReducer(self.policy.flat.numel(), self.policy.flat.device,
                                                           process_group)
            except Exception as e:      # no peer-mapped (symmetric) memory on this system: NCCL does the sum instead
                import sys
                print(f"[tpp_b200] peer-memory all-reduce unavailable ({e!r}); using ncclAllReduce", file=sys.stderr)
                self.optimizer.peer = None
        self.storage.world_size, self.storage.process_group = world_size, process_group

    # ------------------------------------------------------------------------------------------
    def _policy_head(self, obs_slot, storage, env_range=None, slot=0, trunk_only=False, obs_ready=False, heads=True):
        """Policy forward on a rollout slot (or on the env range ``(lo, hi)`` of an image slot, in workspace
        ``slot``) -> head buffer [n, ld_head] (``heads=False``: the embedder's latent pair, see ``_recurrent_head``)."""
        N = storage.num_envs
        if storage.is_image:
            lo, hi = env_range or (0, N)
            n = hi - lo
            c, h, w = storage.obs_shape
            mb = storage.minibatch_buffers(n, *self._obs_buf_args(storage), slot=slot)
            if not obs_ready:          # (obs_ready: the env's step kernel already wrote this slot's rows into mb.obs)
                _lib.call("tpp_frames_to_obs", _lib.ptr(obs_slot[lo:hi]), n, h, w, c, _lib.ptr(mb.obs),
                          _lib.ptr(mb.obs_lo), mb.ld_obs, 1 if mb.raw else 0, _lib.stream_ptr())
                self.n_launches += 1
            return self._fwd(mb.obs, n, x_lo=mb.obs_lo, raw=mb.raw, slot=slot, trunk_only=trunk_only, heads=heads)
        assert env_range is None, "env ranges are implemented for image (row-major frame) slots"
        return self._fwd(obs_slot, N, feature_major_ld=storage.ld, trunk_only=trunk_only, heads=heads)

    def _recurrent_head(self, storage, t):
        """Recurrent policy on slot t (agents/ppo.py:72-81): embedder -> GRU cell on the slot's hidden state, masked by
        the done flags of the PREVIOUS env step (slot 0: the last step of the previous rollout) -> heads.  The new state
        goes to slot t + 1; the bootstrap call (t = T) advances slot T in place -- like the reference, whose
        ``store_last`` keeps the state AFTER the extra predict and starts the next rollout from it (:236-237)."""
        T, N = storage.num_steps, storage.num_envs
        hid = storage.hidden_states_batch
        done_prev = storage.done_carry if t == 0 else storage.done_u8[t - 1]
        pair, ld = self._policy_head(storage.obs_slot(t), storage, heads=False)
        out, ld_out = self.gru.step(pair, ld, hid[t], done_prev, hid[min(t + 1, T)], N)
        return self.engine.head_gemm(out, ld_out, N)

    def _obs_buf_args(self, storage):
        """(row stride, split) of gathered-observation buffers: the TC engine wants TF32 pairs with ld = ceil32(in)."""
        if isinstance(self.engine, MLPEngineTC):
            return self.engine.ld_in, ("raw" if self.engine.raw_pixels and storage.is_image else True)
        return _round4(storage.obs_width), False

    def _policy_sample(self, storage, t, env_range=None, slot=0, obs_ready=False):
        """Rollout step t, policy side: forward on slot t (of an env range) and the action draw into act / logp /
        value of slot t.  MLP policies on the tensor-core engine run the trunk as GEMMs and finish the last layer, the
        heads and the draw in one launch (tpp_mlp_tail_sample)."""
        lo, hi = env_range or (0, storage.num_envs)
        n, eng = hi - lo, self.engine
        act, logp, value = storage.act_i32[t, lo:hi], storage.logp[t, lo:hi], storage.value[t, lo:hi]
        off = int(lo) + getattr(storage, "sample_offset", 0)
        if self.fused_rollout and isinstance(eng, MLPEngineTC):
            raw = bool(storage.is_image and eng.raw_pixels)
            if eng.fused_rollout_ok(raw) and (raw or not storage.is_image):
                if raw and eng.w0_bytes is not None and storage.obs_width <= 768 and storage.obs_width % 4 == 0:
                    # the uint8 frames of the slot ARE the first-layer operand (no TransposeFrame / float copy at all)
                    eng.rollout_fused(storage.obs_slot(t)[lo:hi], n, storage.obs_width, "u8", act, logp, value,
                                      self.sample_seed, self._tick, t, off)
                elif raw:        # pixel rows of the slot: written by the env's step kernel, or converted here
                    c, h, w = storage.obs_shape
                    mb = storage.minibatch_buffers(n, *self._obs_buf_args(storage), slot=slot)
                    if not obs_ready:
                        _lib.call("tpp_frames_to_obs", _lib.ptr(storage.obs_slot(t)[lo:hi]), n, h, w, c,
                                  _lib.ptr(mb.obs), None, mb.ld_obs, 1, _lib.stream_ptr())
                        self.n_launches += 1
                    eng.rollout_fused(mb.obs, n, mb.ld_obs, True, act, logp, value, self.sample_seed, self._tick, t, off)
                else:
                    assert env_range is None
                    eng.rollout_fused(storage.obs_slot(t), n, storage.ld, False, act, logp, value, self.sample_seed,
                                      self._tick, t, off)
                return
        if self.fused_tail and isinstance(eng, MLPEngineTC) and eng.tail_ok():
            h, ldh = self._policy_head(storage.obs_slot(t), storage, env_range=env_range, slot=slot, trunk_only=True,
                                       obs_ready=obs_ready)
            w_off, b_off, fin, fout, relu = eng.layers[-1]
            _lib.call("tpp_mlp_tail_sample", _lib.ptr(h), ldh, fin, eng._p(w_off), eng._p(b_off), fout,
                      1 if relu else 0, eng._p(eng.head_w_off), eng._p(eng.head_b_off), self.n_actions, n, None,
                      eng.ld_head, _lib.ptr(act), _lib.ptr(logp), _lib.ptr(value), self.sample_seed,
                      _lib.ptr(self._tick), int(t), 0, int(lo) + getattr(storage, "sample_offset", 0),
                      _lib.stream_ptr())
            self.n_launches += 1
            return
        if self.recurrent:
            assert env_range is None
            head = self._recurrent_head(storage, t)
        else:
            head = self._policy_head(storage.obs_slot(t), storage, env_range=env_range, slot=slot, obs_ready=obs_ready)
        self._sample(head, n, act, logp, value, t, env_offset=lo + getattr(storage, "sample_offset", 0))

    def _bootstrap_value(self, storage, obs_ready=False):
        """value_batch[T] = V(obs_T) (agents/ppo.py:236-237)."""
        T, N = storage.num_steps, storage.num_envs
        head = self._recurrent_head(storage, T) if self.recurrent \
            else self._policy_head(storage.obs_slot(T), storage, obs_ready=obs_ready)
        storage.value[T, :N] = head[:N, self.n_actions]

    def _fwd(self, x, M, feature_major_ld=None, x_lo=None, raw=False, slot=0, trunk_only=False, heads=True):
        if not heads:
            if isinstance(self.engine, MLPEngineTC):
                return self.engine.forward(x, M, feature_major_ld=feature_major_ld, x_lo=x_lo, need_backward=False,
                                           raw=raw, slot=slot, heads=False)
            return self.engine.forward(x, M, feature_major_ld=feature_major_ld, heads=False)
        if trunk_only:
            return self.engine.forward(x, M, feature_major_ld=feature_major_ld, x_lo=x_lo, need_backward=False, raw=raw,
                                       slot=slot, trunk_only=True)
        if isinstance(self.engine, MLPEngineTC):
            return self.engine.forward(x, M, feature_major_ld=feature_major_ld, x_lo=x_lo, need_backward=False, raw=raw,
                                       slot=slot)
        if slot:
            return self.engine.forward(x, M, feature_major_ld=feature_major_ld, slot=slot)
        return self.engine.forward(x, M, feature_major_ld=feature_major_ld)

    def _sample(self, head, N, act, logp, value, t_offset, greedy=False, env_offset=0):
        _lib.call("tpp_sample_actions", _lib.ptr(head), self.engine.ld_head, N, self.n_actions, _lib.ptr(act),
                  _lib.ptr(logp), _lib.ptr(value), self.sample_seed, _lib.ptr(self._tick), int(t_offset),
                  1 if greedy else 0, int(env_offset), _lib.stream_ptr())
        self.n_launches += 1

    def predict(self, obs, hidden_state, done):
        """Reference signature (agents/ppo.py:72-81): obs [N, *obs_shape] numpy or tensor ->
        (act, log_prob_act, value, hidden_state); numpy in, numpy out."""
        was_numpy = not torch.is_tensor(obs)
        dev = self.policy.flat.device
        x = torch.as_tensor(np.asarray(obs) if was_numpy else obs).to(dev, torch.float32)
        N = x.shape[0]
        x = x.reshape(N, -1).contiguous()
        if self.recurrent:      # (hidden_state, done) -> GRU cell -> heads; returns the next hidden state
            h_np = not torch.is_tensor(hidden_state)
            h = torch.as_tensor(np.asarray(hidden_state) if h_np else hidden_state).to(dev, torch.float32).contiguous()
            d = torch.as_tensor(np.asarray(done) if not torch.is_tensor(done) else done).to(dev).ne(0).to(torch.uint8)
            pair, ld = self._fwd(x, N, heads=False)
            h_next = torch.empty_like(h)
            out, ld_out = self.gru.step(pair, ld, h, d.contiguous(), h_next, N)
            head = self.engine.head_gemm(out, ld_out, N)
            hidden_state = h_next.cpu().numpy() if h_np else h_next
        else:
            head = self._fwd(x, N)
        act = torch.empty(N, dtype=torch.int32, device=dev)
        logp = torch.empty(N, dtype=torch.float32, device=dev)
        value = torch.empty(N, dtype=torch.float32, device=dev)
        self._sample(head, N, act, logp, value, 0)
        _lib.call("tpp_tick_advance", _lib.ptr(self._tick), 1, _lib.stream_ptr())
        if was_numpy:
            return act.cpu().numpy().astype(np.int64), logp.cpu().numpy(), value.cpu().numpy(), hidden_state
        return act, logp, value, hidden_state

    # ------------------------------------------------------------------------------------------
    def optimize(self, defer_summary=False):
        """Reference signature ``optimize() -> summary dict``.  ``defer_summary=True`` returns a closure producing the
        dict instead: everything is enqueued, the device -> host read of the statistics happens when it is called."""
        self._defer_summary = defer_summary
        if self.entropy_scaling == "reward_based":
            mean_rew = np.mean(self.logger.episode_reward_buffer)
            self.entropy_multiplier = 1 - ((mean_rew - self.min_rew) / (self.max_rew - self.min_rew))
        elif self.entropy_scaling == "time_based":
            self.entropy_multiplier = 1 - (self.t / self.total_timesteps)

        st = self.storage
        batch_size = self.n_steps * self.n_envs // self.n_minibatch
        if batch_size < self.mini_batch_size:
            self.mini_batch_size = batch_size
        accum = batch_size / self.mini_batch_size
        mb = self.mini_batch_size
        n_mb = (st.num_steps * st.num_envs) // mb
        if self.recurrent:      # whole trajectories of num_envs // n_mb permuted envs per minibatch (storage.py:93-110)
            envs_per_batch = st.num_envs // max(n_mb, 1)
            if n_mb < 1 or envs_per_batch < 1 or st.num_envs % envs_per_batch:
                raise NotImplementedError("recurrent minibatches: num_envs must split evenly into T*N // mini_batch_size "
                                          "groups of whole trajectories")
            mb, n_mb = st.num_steps * envs_per_batch, st.num_envs // envs_per_batch
        total = n_mb * self.epoch
        dev, A = self.policy.flat.device, self.n_actions
        NS = 4 + 16                                    # doubles per minibatch row of the statistics
        if self._stats is None or self._stats.shape[0] != total:
            self._stats = torch.zeros(total, NS, dtype=torch.float64, device=dev)
        self._stats.zero_()
        # the five loss coefficients live in a small device buffer the loss kernel reads at run time (like Adam's lr), so
        # a captured graph serves every value: entropy_scaling changes the multiplier every call (agents/ppo.py:97-101)
        coef = (self.eps_clip, self.value_coef, self.entropy_coef, float(self.entropy_multiplier), self.x_entropy_coef)
        if getattr(self, "_coef_dev", None) is None:
            self._coef_dev = torch.zeros(8, dtype=torch.float32, device=dev)
            self._coef_pinned = torch.zeros(8, 8, dtype=torch.float32).pin_memory()   # ring: the host may run ahead
            self._coef_host, self._coef_slot = None, 0
        if coef != self._coef_host:
            self._coef_slot = (self._coef_slot + 1) % 8
            self._coef_pinned[self._coef_slot, :5] = torch.tensor(coef, dtype=torch.float32)
            self._coef_dev.copy_(self._coef_pinned[self._coef_slot], non_blocking=True)
            self._coef_host = coef
        cfg = _lib.LossCfg(*coef, A, mb, 0, self._coef_dev.data_ptr())
        engine = self.engine
        fs_vals = []
        is_torch_engine = getattr(engine, "uses_autograd", False)
        step_every = int(accum) if float(accum).is_integer() else 0
        # Gradient-accumulation groups.  The reference runs `accum` minibatches between two optimizer steps
        # (agents/ppo.py:111,173-177) and the weights do not change in between, so G of them (G | accum) share ONE
        # gather / forward / loss / backward pass over G*mb rows: the same sum of per-minibatch mean gradients, with
        # per-minibatch statistics kept (tpp_ppo_loss_fwd_bwd_grouped).  Kernels of 8192-row minibatches are
        # launch- and prologue-bound on 148 SMs; G*mb rows fill the machine (profiles/README.md).
        G = self.group_size = self._group_size(step_every, n_mb, mb, engine)
        rows, n_grp = G * mb, n_mb // G
        buf = st.minibatch_buffers(rows, *self._obs_buf_args(st))
        self.policy.train()
        # fixed staging buffers so that one group (gather -> forward -> loss -> backward) is a replayable graph
        if getattr(self, "_idx_cur", None) is None or self._idx_cur.numel() != rows:
            self._idx_cur = torch.zeros(rows, dtype=torch.int64, device=dev)
            self._stats_cur = torch.zeros(G, NS, dtype=torch.float64, device=dev)
            self._pbar_cur = torch.zeros(16, dtype=torch.float32, device=dev)
        if hasattr(engine, "refresh_weights"):
            engine.refresh_weights()

        def group_body(idx_rows=None, stats_rows=None):
            s = _lib.stream_ptr()          # evaluated here: under graph capture the current stream is the capture stream
            idx_rows = self._idx_cur if idx_rows is None else idx_rows
            stats_rows = self._stats_cur if stats_rows is None else stats_rows
            stats_rows.zero_()
            st.gather(idx_rows, buf)
            if is_torch_engine or isinstance(engine, ImpalaEngineTC):
                head = engine.forward(buf.obs, rows, train=True)
            elif buf.obs_lo is not None or buf.raw:
                head = engine.forward(buf.obs, rows, x_lo=buf.obs_lo, raw=buf.raw)
            else:
                head = engine.forward(buf.obs, rows)
            ws_dhead = engine._workspace(rows).dhead if not is_torch_engine else self._dhead(rows)
            if G > 1:
                _lib.call("tpp_ppo_loss_fwd_bwd_grouped", C.byref(cfg), G, _lib.ptr(head), engine.ld_head,
                          _lib.ptr(buf.act), _lib.ptr(buf.logp), _lib.ptr(buf.value), _lib.ptr(buf.ret),
                          _lib.ptr(buf.adv), _lib.ptr(ws_dhead), _lib.ptr(stats_rows), NS, s)
            else:
                pbar = None
                if self.x_entropy_coef != 0.0:
                    self._pbar_cur.zero_()
                    _lib.call("tpp_ppo_pbar", _lib.ptr(head), engine.ld_head, mb, A, _lib.ptr(self._pbar_cur), s)
                    pbar = self._pbar_cur
                    self.n_launches += 1
                _lib.call("tpp_ppo_loss_fwd_bwd", C.byref(cfg), _lib.ptr(head), engine.ld_head, _lib.ptr(buf.act),
                          _lib.ptr(buf.logp), _lib.ptr(buf.value), _lib.ptr(buf.ret), _lib.ptr(buf.adv),
                          _lib.ptr(pbar), _lib.ptr(ws_dhead), _lib.ptr(stats_rows), s)
            self.n_launches += 1
            if is_torch_engine or isinstance(engine, ImpalaEngineTC):
                engine.backward(ws_dhead, rows, self.fs_coef)
            else:
                engine.backward(ws_dhead, rows)

        graph_key = (mb, G, self.x_entropy_coef != 0.0, id(st))
        use_graph = self.use_cuda_graph and not is_torch_engine
        graphs = self.__dict__.setdefault("_mb_graphs", {})
        k = 0
        # Whole-epoch graph (MLP engines): every group of an epoch -- gather, forward, loss, backward --
        # and the optimizer steps + weight re-splits at their fixed positions are ONE graph replay per epoch; the
        # epoch's permutation is uploaded into a static index buffer first.  Removes the per-group graph launches, index
        # copies and stats copies of an iteration from the host's critical path.
        epoch_graph = (use_graph and self.use_epoch_graph and (self.world_size == 1 or self.graph_allreduce)
                       and step_every > 0
                       and n_mb % step_every == 0
                       and isinstance(engine, (MLPEngine, MLPEngineTC)) and self.x_entropy_coef == 0.0
                       and not self.recurrent)
        if epoch_graph:
            if getattr(self, "_epoch_idx", None) is None or self._epoch_idx.shape != (n_mb, mb):
                self._epoch_idx = torch.zeros(n_mb, mb, dtype=torch.int64, device=dev)
                self._epoch_stats = torch.zeros(n_mb, NS, dtype=torch.float64, device=dev)
            egraphs = self.__dict__.setdefault("_epoch_graphs", {})
            ekey = (mb, n_mb, G, step_every, id(st))
            idx_g, stats_g = self._epoch_idx.view(n_grp, rows), self._epoch_stats.view(n_grp, G, NS)

            def epoch_body():
                for i in range(n_grp):
                    group_body(idx_g[i], stats_g[i])
                    if ((i + 1) * G) % step_every == 0:
                        if self.world_size > 1 and self.optimizer.peer is None:   # NCCL all-reduce captured in the graph
                            parallel.allreduce_gradients_(self.policy.flat_grad, self.process_group)
                        views = engine.weight_views() if hasattr(engine, "weight_views") else None
                        self.optimizer.launch(views)
                        if views is None and hasattr(engine, "refresh_weights"):
                            engine.refresh_weights()

            def counts():
                return (self.n_launches, self.engine.n_launches, self.storage.n_launches, self.optimizer.n_launches)

            # Index upload: epoch e + 1's permutation is drawn on the host and copied up on a side stream while epoch e's
            # graph runs (double-buffered staging), unless the caller supplies device-resident indices itself
            # (``storage.epoch_indices`` replaced, as bench.py's device-timed leg does).
            pipelined = "epoch_indices" not in st.__dict__
            if pipelined:
                main = torch.cuda.current_stream()
                if getattr(self, "_h2d_stream", None) is None:
                    self._h2d_stream = torch.cuda.Stream()
                if getattr(self, "_idx_stage", None) is None or self._idx_stage[0].shape != (n_mb, mb):
                    self._idx_stage = [torch.zeros(n_mb, mb, dtype=torch.int32, device=dev) for _ in range(2)]
                    self._stage_free = [None, None]

                def upload(e):
                    host, slot_state = st.epoch_perm_pinned(mb, e)
                    side = self._h2d_stream
                    if self._stage_free[e % 2] is not None:
                        side.wait_event(self._stage_free[e % 2])      # main has consumed this staging buffer
                    with torch.cuda.stream(side):
                        self._idx_stage[e % 2].copy_(host, non_blocking=True)
                        done = torch.cuda.Event()
                        done.record(side)
                    slot_state[1] = done
                    return done
                pending = upload(0)
            for e in range(self.epoch):
                if pipelined:
                    main.wait_event(pending)
                    self._epoch_idx.copy_(self._idx_stage[e % 2])
                    self._stage_free[e % 2] = torch.cuda.Event()
                    self._stage_free[e % 2].record(main)
                else:
                    self._epoch_idx.copy_(st.epoch_indices(mb))
                self.optimizer._sync_lr()
                entry = egraphs.get(ekey)
                if entry is None:                      # first epoch ever: eager (allocates workspaces)
                    epoch_body()
                    egraphs[ekey] = "warm"
                else:
                    if entry == "warm":
                        g = torch.cuda.CUDAGraph()
                        torch.cuda.synchronize()
                        c0 = counts()
                        with torch.cuda.graph(g):
                            epoch_body()
                        delta = tuple(b - a for a, b in zip(c0, counts()))
                        self.n_launches, self.engine.n_launches, self.storage.n_launches, \
                            self.optimizer.n_launches = c0
                        entry = egraphs[ekey] = (g, delta)
                    entry[0].replay()
                    self.n_launches += entry[1][0]
                    self.engine.n_launches += entry[1][1]
                    self.storage.n_launches += entry[1][2]
                    self.optimizer.n_launches += entry[1][3]
                if pipelined and e + 1 < self.epoch:
                    pending = upload(e + 1)             # host draw + H2D overlap the epoch just launched
                self._stats[k:k + n_mb].copy_(self._epoch_stats)
                k += n_mb
            return self._summary(fs_vals)
        for _ in range(self.epoch):
            idx = (st.epoch_indices_recurrent(self.mini_batch_size) if self.recurrent
                   else st.epoch_indices(mb)).view(n_grp, rows)
            for i in range(n_grp):
                self._idx_cur.copy_(idx[i])
                entry = graphs.get(graph_key) if use_graph else None
                if entry is None or entry == "warm":
                    if use_graph and entry == "warm":                  # second group: capture, then replay
                        g = torch.cuda.CUDAGraph()
                        torch.cuda.synchronize()
                        c0 = self._launch_count()
                        with torch.cuda.graph(g):
                            group_body()
                        n_captured = sum(self._launch_count()) - sum(c0)
                        # kernels recorded during capture did not execute: count them per replay instead
                        self.n_launches, self.engine.n_launches, self.storage.n_launches = c0
                        entry = graphs[graph_key] = (g, n_captured)
                    else:                                              # first group (or graphs off): eager
                        group_body()
                        if use_graph:
                            graphs[graph_key] = "warm"
                if isinstance(entry, tuple):
                    entry[0].replay()
                    self.n_launches += entry[1]
                if getattr(engine, "last_fs", None) is not None:
                    fs_vals.append(engine.last_fs.detach().clone())
                self._stats[k:k + G].copy_(self._stats_cur)
                k += G
                if k % accum == 0:                 # k minibatches done this call (reference: cnt % accum, :173)
                    if self.world_size > 1 and self.optimizer.peer is None:
                        parallel.allreduce_gradients_(self.policy.flat_grad, self.process_group)
                    views = engine.weight_views() if hasattr(engine, "weight_views") else None
                    self.optimizer.step(views)
                    if views is None and hasattr(engine, "refresh_weights"):
                        engine.refresh_weights()
        return self._summary(fs_vals)

    def _group_size(self, step_every, n_mb, mb, engine):
        """Minibatches sharing one forward/backward pass (see optimize): the largest divisor of the accumulation
        count that the ``fuse_accum`` setting ("auto" = all of them, or an int cap; 1 / False = off) and the row
        budget allow."""
        want = self.__dict__.get("fuse_accum", "auto")
        if want in (False, None, 0, 1) or step_every <= 1 or n_mb % step_every != 0 or mb % 256 != 0:
            return 1
        if not isinstance(engine, (MLPEngine, MLPEngineTC)) or self.x_entropy_coef != 0.0:
            return 1
        cap = step_every if want in ("auto", True) else max(1, int(want))
        cap = min(cap, max(1, self.max_group_rows // mb))
        return max(g for g in range(1, step_every + 1) if step_every % g == 0 and g <= cap)

    def _launch_count(self):
        return (self.n_launches, self.engine.n_launches, self.storage.n_launches)

    def _dhead(self, mb):
        if getattr(self, "_dhead_buf", None) is None or self._dhead_buf.shape[0] != mb:
            self._dhead_buf = torch.zeros(mb, self.engine.ld_head, dtype=torch.float32, device=self.policy.flat.device)
        return self._dhead_buf

    def _summary(self, fs_vals):
        if getattr(self, "_defer_summary", False):
            if getattr(self, "_stats_host", None) is None or self._stats_host.shape != self._stats.shape:
                self._stats_host = torch.empty(self._stats.shape, dtype=torch.float64).pin_memory()
            self._stats_host.copy_(self._stats, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()

            def finish():
                ev.synchronize()
                return self._summary_from(self._stats_host.numpy(), fs_vals)
            return finish
        return self._summary_from(self._stats.cpu().numpy(), fs_vals)

    def _summary_from(self, S, fs_vals):
        """One device->host read of the accumulated sums, then the reference's nine summary keys in order
        (agents/ppo.py:199-207).  Loss/pi and Loss/v keep the reference's sign quirk (negated)."""
        A = self.n_actions
        B = S[:, 3]
        pi_loss = -S[:, 0] / B
        v_loss = 0.5 * S[:, 1] / B
        ent = S[:, 2] / B
        pbar = S[:, 4:4 + A] / B[:, None]
        marg = -(pbar * np.log(pbar)).sum(1)
        x_ent = marg - ent
        fs = np.array([float(v) for v in fs_vals]) if fs_vals else None
        total = (pi_loss + self.value_coef * v_loss - self.entropy_coef * ent * self.entropy_multiplier
                 - self.x_entropy_coef * x_ent)
        if fs is not None:
            total = total + self.fs_coef * fs
        self.last_stats = dict(pi_loss=pi_loss, value_loss=v_loss, entropy=ent, x_entropy=x_ent, total=total)
        nan = float("nan")
        return {"Loss/pi": float(np.mean(-pi_loss)), "Loss/v": float(np.mean(-v_loss)),
                "Loss/entropy": float(np.mean(ent)), "Loss/x_entropy": float(np.mean(x_ent)),
                "Loss/atn_entropy": nan, "Loss/atn_entropy2": nan, "Loss/sparsity": nan,
                "Loss/feature_sparsity": float(np.mean(fs)) if fs is not None else nan,
                "Loss/total": float(np.mean(total))}

    # ------------------------------------------------------------------------------------------
    # Rollout
    # ------------------------------------------------------------------------------------------
    def _device_env(self, env):
        return hasattr(env, "rollout_step")

    def _env_ranges(self, env, storage):
        """Env ranges that step through the rollout as independent kernel chains on their own streams.  One step of
        a few thousand envs is a chain of ~10 dependent launches of 30-130 CTAs each -- latency-bound, most SMs idle;
        envs are independent, so ranges overlap each other's launches.  ``rollout_chains``: "auto" (ranges of >= 1024
        envs, at most 4), an int, or 1 = off."""
        N = storage.num_envs
        want = self.rollout_chains
        if not getattr(env, "supports_env_ranges", False) or not storage.is_image \
                or not isinstance(self.engine, (MLPEngine, MLPEngineTC)):
            return [(0, N)]
        C = min(4, N // 1024) if want == "auto" else int(want)
        C = max(1, min(C, N // 128))
        step = -(-N // C // 128) * 128          # ranges start on multiples of 128 envs (GEMM row tiles)
        return [(lo, min(N, lo + step)) for lo in range(0, N, step)]

    def _rollout_steps(self, env, storage):
        T, N = storage.num_steps, storage.num_envs
        if hasattr(self.engine, "refresh_weights"):
            self.engine.refresh_weights()     # always part of the (captured) rollout: weights changed since last time
        if self.gru is not None:
            self.gru.refresh_weights()
        ranges = self._env_ranges(env, storage)
        # envs that can emit the policy's next input rows from their step kernel (Box-World, raw-pixel first layer)
        # save the frames -> obs launch of every step: only slot 0 is converted here
        u8_direct = (self.fused_rollout and isinstance(self.engine, MLPEngineTC) and self.engine.raw_pixels
                     and self.engine.fused_rollout_ok(True) and getattr(self.engine, "w0_bytes", None) is not None
                     and storage.obs_width <= 768 and storage.obs_width % 4 == 0)
        fold = (len(ranges) == 1 and getattr(env, "emits_policy_obs", False) and storage.is_image
                and isinstance(self.engine, MLPEngineTC) and self.engine.raw_pixels and not u8_direct
                and not self.recurrent)
        if fold:
            mb = storage.minibatch_buffers(N, *self._obs_buf_args(storage))
            assert mb.raw and mb.obs_lo is None
            for t in range(T):
                self._policy_sample(storage, t, obs_ready=t > 0)
                env.rollout_step(storage, t, obs_out=mb.obs)
        elif len(ranges) == 1:
            for t in range(T):
                self._policy_sample(storage, t)
                env.rollout_step(storage, t)
        else:
            self._rollout_chains(env, storage, ranges)
        if hasattr(env, "finish_rollout"):
            env.finish_rollout(storage)
        self._bootstrap_value(storage, obs_ready=fold)
        _lib.call("tpp_tick_advance", _lib.ptr(self._tick), T, _lib.stream_ptr())
        env.advance_tick(T)
        self.n_launches += 2

    def _rollout_chains(self, env, storage, ranges):
        """T steps of every env range, one stream per range (forked from / joined into the current stream, so the
        whole thing is capturable as one CUDA graph with parallel branches).  The only cross-range dependency is the
        env's sequential level-seed counter: the env transitions of one step run in ascending env order, and step
        t + 1 of the first range follows step t of the last (events)."""
        T = storage.num_steps
        main = torch.cuda.current_stream()
        if getattr(self, "_chain_streams", None) is None or len(self._chain_streams) != len(ranges):
            self._chain_streams = [torch.cuda.Stream() for _ in ranges]
        streams = self._chain_streams
        fork = torch.cuda.Event()
        fork.record(main)
        for s in streams:
            s.wait_event(fork)
        prev = None                                   # event after the previous env transition (global order)
        for t in range(T):
            for c, (lo, hi) in enumerate(ranges):
                with torch.cuda.stream(streams[c]):
                    self._policy_sample(storage, t, env_range=(lo, hi), slot=1 + c)
                    if prev is not None:
                        streams[c].wait_event(prev)
                    env.rollout_step(storage, t, env_range=(lo, hi))
                    prev = torch.cuda.Event()
                    prev.record(streams[c])
        for s in streams:
            join = torch.cuda.Event()
            join.record(s)
            main.wait_event(join)

    def collect_rollout(self, env=None, storage=None):
        """T fused steps on the device (policy -> sample -> env) followed by the bootstrap value."""
        env, storage = env or self.env, storage or self.storage
        key = (id(env), id(storage))
        if not self.use_cuda_graph:
            return self._rollout_steps(env, storage)
        graphs = self.__dict__.setdefault("_graphs", {})
        if key not in graphs:
            self._rollout_steps(env, storage)            # eager warm-up (allocates workspaces, loads modules)
            graphs[key] = None
            return
        if graphs[key] is None:
            g = torch.cuda.CUDAGraph()
            torch.cuda.synchronize()
            launches = (self.n_launches, self.engine.n_launches)
            with torch.cuda.graph(g):
                self._rollout_steps(env, storage)
            graphs[key] = (g, self.n_launches - launches[0], self.engine.n_launches - launches[1])
        g, d_self, d_eng = graphs[key]
        g.replay()
        self.n_launches += d_self
        self.engine.n_launches += d_eng

    def _carry_over(self, storage):
        """Slot T of the finished rollout is slot 0 of the next one."""
        T = storage.num_steps
        storage.obs_slot(0).copy_(storage.obs_slot(T))
        if self.recurrent:      # hidden state and done flags travel with the observation (agents/ppo.py:219-237)
            storage.hidden_states_batch[0].copy_(storage.hidden_states_batch[T])
            storage.done_carry.copy_(storage.done_u8[T - 1])

    def train(self, num_timesteps):
        self.total_timesteps = num_timesteps
        save_every = num_timesteps // self.num_checkpoints if self.num_checkpoints else num_timesteps + 1
        checkpoints = sorted((i + 1) * save_every for i in range(max(self.num_checkpoints, 1)))
        checkpoint_cnt = 0
        if not self._device_env(self.env):
            return self._train_host_env(num_timesteps, checkpoints)
        self.env.reset_rollout(self.storage)
        if self.env_valid is not None:
            self.env_valid.reset_rollout(self.storage_valid)

        def rollouts():
            self.policy.eval()
            self.collect_rollout(self.env, self.storage)
            self.storage.compute_estimates(self.gamma, self.lmbda, self.use_gae, self.normalize_adv)
            if self.env_valid is not None:
                self.collect_rollout(self.env_valid, self.storage_valid)
                self.storage_valid.compute_estimates(self.gamma, self.lmbda, self.use_gae, self.normalize_adv)

        # Software pipeline of the host side: iteration i's update is enqueued, its statistics / log batches start their
        # device -> host copies, and iteration i + 1's rollout (which only needs the updated weights, in stream order) is
        # enqueued BEFORE the host waits for those copies and runs the logger -- the GPU never idles behind host work.
        rollouts()
        while self.t < num_timesteps:
            summary_fn = self.optimize(defer_summary=True)
            logs = self._snapshot_logs()
            self.t += self.n_steps * self.n_envs
            self._carry_over(self.storage)
            if self.env_valid is not None:
                self._carry_over(self.storage_valid)
            # reference order (agents/ppo.py:257-276): t advances, the learning rate of the NEXT update is set, then the
            # checkpoint is written -- so a checkpoint's optimizer state carries the adjusted rate
            self.optimizer, lr = self.adjust_lr(self.optimizer, self.learning_rate, self.t, num_timesteps)
            save = self.num_checkpoints and checkpoint_cnt < len(checkpoints) and self.t > checkpoints[checkpoint_cnt]
            if save:                                   # (weights of iteration i: before the next update exists)
                self.save_checkpoint()
                checkpoint_cnt += 1
            if self.t < num_timesteps:
                rollouts()
            self._log(summary_fn(), lr, logs)
        self.env.close()
        if self.env_valid is not None:
            self.env_valid.close()

    def _snapshot_logs(self):
        """Start the device -> host copies the logger needs for the rollout just consumed.  A logger with
        ``feed_episodes`` (this package's) gets the device-side episode records (a few hundred bytes); any other
        object with the reference's ``feed`` gets the [T, N] reward / done batches."""
        if self.logger is None:
            return None
        sv = self.storage_valid
        if hasattr(self.logger, "feed_episodes"):
            return ("episodes", self.storage.snapshot_episodes(), sv.snapshot_episodes() if sv is not None else None)
        return ("batches", self.storage.snapshot_log_data(), sv.snapshot_log_data() if sv is not None else None)

    def _log(self, summary, lr, snapshots=None):
        if self.logger is not None:
            if snapshots is None:
                snapshots = self._snapshot_logs()
            kind, tr, va = snapshots
            if kind == "episodes":
                nan = float("nan")
                self.logger.feed_episodes(self.n_steps, tr(), nan, va() if va is not None else None, nan)
            else:
                rew_batch, done_batch, tar = tr()
                rew_v, done_v, tar_v = va() if va is not None else (None, None, None)
                self.logger.feed(rew_batch, done_batch, tar, rew_v, done_v, tar_v)
            self.logger.dump(summary, lr)

    def save_checkpoint(self):
        """Same file name and keys as the reference (agents/ppo.py:271-276)."""
        if self.logger is None or getattr(self.logger, "logdir", None) is None:
            return
        torch.save({"model_state_dict": self.policy.state_dict(), "optimizer_state_dict": self.optimizer.state_dict()},
                   self.logger.logdir + "/model_" + str(self.t) + ".pth")

    def _host_step_device(self, st, t, N):
        """Device side of one host-env step (frames -> policy forward -> Philox sampling into slot t, then the start of
        the actions' device -> host copy).  The ~40 launches are captured once per slot (second visit) and replayed
        afterwards: the step is launch-bound."""
        def body():
            head = self._recurrent_head(st, t) if self.recurrent else self._policy_head(st.obs_slot(t), st)
            self._sample(head, N, st.act_i32[t], st.logp[t], st.value[t], t,
                         env_offset=getattr(st, "sample_offset", 0))

        graphs = self.__dict__.setdefault("_host_graphs", {})
        key = (id(st), t)
        entry = graphs.get(key)
        if not self.use_cuda_graph or getattr(self.engine, "uses_autograd", False):
            body()
        elif entry is None:                    # first visit: eager (allocates the workspaces)
            body()
            graphs[key] = "warm"
        else:
            if entry == "warm":
                g = torch.cuda.CUDAGraph()
                torch.cuda.synchronize()
                c0 = (self.n_launches, self.engine.n_launches)
                with torch.cuda.graph(g):
                    body()
                entry = graphs[key] = (g, self.n_launches - c0[0], self.engine.n_launches - c0[1])
                self.n_launches, self.engine.n_launches = c0
            entry[0].replay()
            self.n_launches += entry[1]
            self.engine.n_launches += entry[2]
        st.start_action_fetch(t)

    def _train_host_env(self, num_timesteps, checkpoints):
        """Host-stepped envs (Procgen or any numpy VecEnv; reference loop agents/ppo.py:216-279).  Per step: the
        observation is staged into rollout slot t (double-buffered pinned H2D; uint8 frames stay uint8), the policy
        forward and the action draw run on the device on that slot, only the N actions come back to the host (pinned,
        event wait -- no stream synchronisation) for ``env.step``; rewards / dones collect in pinned host rows and go up
        once per rollout.  The validation env steps in the same loop like the reference's (:228-252): its device work
        overlaps the training env's host step.  GAE and the update never leave the device."""
        checkpoint_cnt = 0
        N, T = self.n_envs, self.n_steps
        pairs = [(self.env, self.storage)]
        if self.env_valid is not None:
            pairs.append((self.env_valid, self.storage_valid))
        raw = [bool(getattr(e, "stages_raw_frames", False)) for e, _ in pairs]
        obs = [e.host_reset() if r else e.reset() for (e, _), r in zip(pairs, raw)]
        while self.t < num_timesteps:
            self.policy.eval()
            if hasattr(self.engine, "refresh_weights"):
                self.engine.refresh_weights()
            if self.gru is not None:
                self.gru.refresh_weights()
            for t in range(T):
                for k, (e, st) in enumerate(pairs):
                    st.stage_obs(t, obs[k])
                    self._host_step_device(st, t, N)
                for k, (e, st) in enumerate(pairs):
                    act = st.finish_action_fetch()                 # the step's only device -> host read
                    if raw[k]:
                        obs[k], rew, done, info = e.host_step(act)
                        st.stage_step(t, rew, done, raw_rew=rew, upload_done=self.recurrent)
                    else:
                        obs[k], rew, done, info = e.step(act)
                        st.stage_step(t, rew, done, info=info, upload_done=self.recurrent)
            for k, (e, st) in enumerate(pairs):
                st.stage_obs(T, obs[k])
                st.flush_steps()
                if hasattr(e, "finish_rollout"):
                    e.finish_rollout(st)
                self._bootstrap_value(st)
                st.compute_estimates(self.gamma, self.lmbda, self.use_gae, self.normalize_adv)
                if self.recurrent:      # the next rollout's slot 0 (stage_obs refills the observation itself)
                    st.hidden_states_batch[0].copy_(st.hidden_states_batch[T])
                    st.done_carry.copy_(st.done_u8[T - 1])
            _lib.call("tpp_tick_advance", _lib.ptr(self._tick), T, _lib.stream_ptr())
            summary = self.optimize()
            self.t += T * N
            self.optimizer, lr = self.adjust_lr(self.optimizer, self.learning_rate, self.t, num_timesteps)
            self._log(summary, lr)
            if self.num_checkpoints and checkpoint_cnt < len(checkpoints) and self.t > checkpoints[checkpoint_cnt]:
                self.save_checkpoint()
                checkpoint_cnt += 1
        for e, _ in pairs:
            e.close()


def _round4(x):
    return (x + 3) // 4 * 4
