#!/usr/bin/env python
"""Benchmark of the PPO rollout-and-update hot path (BASELINE.json metric: PPO env-steps/sec, rollout+GAE+update).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload boxworld|cartpole|procgen|procgen_c5]

One "step" = one full PPO iteration: T fused rollout steps over the rank's envs (policy forward -> Philox action
sampling -> env step kernel writing into the rollout), bootstrap value, GAE + advantage normalisation, and
`epoch` x minibatch updates (device gather -> policy fwd -> fused loss fwd+bwd -> policy bwd -> clip+Adam).

Hyperparameters come UNCHANGED from the reference's YAML sets (hyperparams/procgen/config.yml, read through
tpp_b200.helper_local.get_hyperparams from $TPP_CONFIG_YML or the extracted copy tests/golden/config_subset.yml);
only what BASELINE.json's configs name differently is overridden (n_envs).

Default workload = BASELINE.json configs[1] (C2): boxworld_env_vec PPO, n_envs=4096 IN TOTAL, env-sharded over the N
GPUs (`scaling: strong`; the same run also reports weak scaling, 4096 envs per GPU, under `weak_scaling`), YAML set
`boxworld-impala` (config.yml:575-601: T=256, 3 epochs, n_minibatch 8, mini_batch_size 8192, grid 12, goal 5, 3x3
distractors), 500-level bank, with an MLP policy on the flattened 3x14x14 frame (the reference has no working
Box-World policy; choice documented in DESIGN.md).  At N=1 the JSON line also carries a `configs` block:
C1 (`cartpole` set, 256 envs x 256 steps: device-timed, end-to-end and the CPU arm at the SAME 256 x 256), the C3
env-step sweep (`kernel_rooflines`), C4 (`hard-500` set, IMPALA-CNN, 64 envs, synthetic host engine) and C5 (the
`hard-500` set as it is: 256 envs, minibatch 8192 -- the maze_aisc / heist_aisc_many_chests shape).

`value` = device-timed iterations with inputs resident (minibatch permutations pre-uploaded); `e2e` = the same
iterations through the public API `PPO.train()` with this package's `Logger`: per-epoch index upload from pinned host
memory, device->host reads of the loss summary and of the logger's episode records.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIG_YML = os.environ.get("TPP_CONFIG_YML") or os.path.join(ROOT, "tests", "golden", "config_subset.yml")
PPO_KEYS = ("n_steps", "n_envs", "epoch", "n_minibatch", "mini_batch_size", "gamma", "lmbda", "learning_rate",
            "grad_clip_norm", "eps_clip", "value_coef", "entropy_coef")
METRIC = "PPO env-steps/sec (rollout+GAE+update)"


def workload_hp(name):
    """YAML set of the workload + the overrides BASELINE.json's configs ask for."""
    import yaml
    with open(CONFIG_YML) as f:
        sets = yaml.safe_load(f)
    if name == "boxworld":          # configs[1]: n_envs=4096; MLP policy on the flattened frame
        return dict(sets["boxworld-impala"], n_envs=4096, max_steps=1000, depth=4, mid_weight=256, latent_size=64)
    if name == "cartpole":          # configs[0]: the set as it is
        return dict(sets["cartpole"])
    if name == "procgen":           # configs[3]: hard-500 with 64 envs per GPU (n_minibatch: PPO's default, 8)
        return dict(sets["hard-500"], n_envs=64, n_minibatch=8)
    if name == "procgen_c5":        # configs[4]: the hard-500 set as it is (256 envs, minibatch 8192) -- the shape the
        return dict(sets["hard-500"], n_minibatch=8)      # maze_aisc / heist_aisc_many_chests 200M-step runs use
    raise KeyError(name)


# bounded CPU samples of the same workloads: (n_envs, n_steps) of one CPU "step"; everything else is the GPU arm's
CPU_SAMPLE = {"boxworld": (4096, 16), "cartpole": (256, 256), "procgen": (64, 32), "procgen_c5": (256, 8)}


def workload_string(name, hp):
    mb = min(hp["mini_batch_size"], hp["n_steps"] * hp["n_envs"] // hp["n_minibatch"])
    pol = {"boxworld": "MLP policy 588-256-256-256-64 on the flattened 3x14x14 frame",
           "cartpole": "MLP policy 9-256-256-256-64", "procgen": "IMPALA-CNN policy, synthetic 64x64x3 uint8 frames",
           "procgen_c5": "IMPALA-CNN policy, synthetic 64x64x3 uint8 frames"}[name]
    return (f"{name} PPO: n_envs={hp['n_envs']}, n_steps={hp['n_steps']}, epoch={hp['epoch']}, "
            f"n_minibatch={hp['n_minibatch']}, mini_batch_size={mb}, {pol}")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], bf16=d["bf16_tflops"], bf16_sustained=d.get("bf16_tflops_sustained"),
                    source="measured")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
# Our arm
# --------------------------------------------------------------------------------------------------

def build_agent(name, hp, rank, device, **extra):
    from tpp_b200.agents.ppo import PPO
    from tpp_b200.common.model import MLPModel
    from tpp_b200.common.policy import CategoricalPolicy
    from tpp_b200.common.storage import Storage
    N, T = hp["n_envs"], hp["n_steps"]
    if name == "boxworld":
        from tpp_b200.boxworld.box_world_env_vec import create_bw_env

        class A:
            seed, num_levels = 6033 + 1000 * rank, 500
        env = create_bw_env(A, dict(hp, device=device))
    else:
        from tpp_b200.discrete_env.cartpole_pre_vec import create_cartpole

        class A:
            seed = 6033 + rank
        env = create_cartpole(A, dict(hp, device=device))
    obs_shape = env.observation_space.shape
    in_dim = int(np.prod(obs_shape))
    torch.manual_seed(6033)                      # identical initial weights on every rank
    pol = CategoricalPolicy(MLPModel(in_dim, hp["depth"], hp["mid_weight"], hp["latent_size"]), False,
                            env.action_space.n).to(device).flatten_()
    st = Storage(obs_shape, hp["latent_size"], T, N, device)
    agent = PPO(env, pol, None, st, device, 0, **{k: hp[k] for k in PPO_KEYS}, sample_seed=17, **extra)
    return agent, in_dim


def time_kernel(fn, iters=20, warm=3):
    """Average device time of one `fn()` (one or more launches on the current stream): `iters` calls are captured in
    a CUDA graph and the replay is bracketed by CUDA events, so the number contains no Python / ctypes launch cost
    (which is ~10 us per call and would otherwise bound every kernel shorter than that)."""
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(iters):
                fn()
        g.replay()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        g.replay()
        b.record()
        torch.cuda.synchronize()
    return a.elapsed_time(b) / (2 * iters) * 1e-3     # seconds per call


def kernel_rooflines(pk, device, sweep=True):
    """Live CUDA-event timings of the HBM-bound kernels, achieved = algorithmic bytes (SURVEY 8d per-unit figures,
    DESIGN.md section 4) / time.  Env steps ping-pong between two rollout slots; at N = 2^22 one slot is 75-235 MB,
    i.e. the working set exceeds the 126 MB L2 (smaller N of the C3 sweep are L2-resident and say so)."""
    from tpp_b200.boxworld.box_world_env_vec import BoxWorldVec
    from tpp_b200.common.storage import Storage
    from tpp_b200.discrete_env.acrobot_pre_vec import AcrobotVecEnv
    from tpp_b200.discrete_env.cartpole_pre_vec import CartPoleVecEnv
    from tpp_b200.discrete_env.cartpole_swing_pre_vec import CartPoleSwingVecEnv
    from tpp_b200.discrete_env.lunar_lander_pre_vec import LunarLanderVecEnv
    from tpp_b200.discrete_env.mountain_car_pre_vec import MountainCarVecEnv
    out = []
    fams = ((CartPoleVecEnv, 89), (CartPoleSwingVecEnv, 89), (MountainCarVecEnv, 57), (AcrobotVecEnv, 137),
            (LunarLanderVecEnv, 81))
    for cls, bytes_per_step in fams:
        sizes = [1 << 22]
        if sweep and cls in (AcrobotVecEnv, MountainCarVecEnv, LunarLanderVecEnv):     # BASELINE configs[2] (C3)
            sizes = [1 << 16, 1 << 18, 1 << 20, 1 << 22]
        for N in sizes:
            env = cls(n_envs=N, seed=1, device=device)
            act = torch.randint(0, env.n_actions, (N,), device=device, dtype=torch.int32)
            state = {"cur": 0}

            def step():
                c = state["cur"]
                env.step_into(env._slots[c], env._slots[c ^ 1], act, env._rew, env._done)
                state["cur"] = c ^ 1
            dt = time_kernel(step, iters=50)
            gbs = bytes_per_step * N / dt / 1e9
            ws = 2 * env.n_obs * 4 * N / 1e6
            out.append(dict(kernel=f"env_step_{env.family}", n_envs=N, bytes_per_unit=bytes_per_step, bound="hbm",
                            achieved=round(gbs, 1), peak=pk["hbm"], unit="GB/s", frac=round(gbs / pk["hbm"], 4),
                            env_steps_per_s=round(N / dt, 1), us=round(dt * 1e6, 2), working_set_mb=round(ws, 1),
                            l2_resident=bool(ws < 126)))
            del env
    # Box-World step (+ in-order level replacement + frame emit into the rollout slot), uint8 rollout contract:
    # read frame 588 + write frame 588 + ~57 B of cells / meta / reward / done per env-step
    for N in (4096, 1 << 18):
        bw = BoxWorldVec(N, 12, 5, 3, 3, max_steps=1000, start_seed=6033, n_levels=500, device=device)
        frames = torch.zeros(2, N, 14, 14, 3, dtype=torch.uint8, device=device)
        act = torch.randint(0, 4, (N,), device=device, dtype=torch.int32)
        k = {"i": 0}

        def bstep():
            k["i"] ^= 1
            bw.step_device(act, frame_out=frames[k["i"]])
        dt = time_kernel(bstep, iters=50)
        gbs = 1233 * N / dt / 1e9
        out.append(dict(kernel="boxworld_step+reset (n=12)", n_envs=N, bytes_per_unit=1233, bound="hbm",
                        achieved=round(gbs, 1), peak=pk["hbm"], unit="GB/s", frac=round(gbs / pk["hbm"], 4),
                        env_steps_per_s=round(N / dt, 1), us=round(dt * 1e6, 2)))
        del bw, frames
    for T, Ng in ((256, 4096), (256, 1 << 16), (64, 1 << 20)):
        st = Storage((1,), 1, T, Ng, device)
        st.rew.normal_(); st.value.normal_()
        st.done_u8.copy_((torch.rand(T, st.ld, device=device) < 0.02).to(torch.uint8))
        dt = time_kernel(lambda: st.compute_estimates(0.99, 0.95, True, True), iters=10)
        gbs = 25 * T * Ng / dt / 1e9
        out.append(dict(kernel="gae_scan+adv_normalize", T=T, n_envs=Ng, bytes_per_unit=25, bound="hbm",
                        achieved=round(gbs, 1), peak=pk["hbm"], unit="GB/s", frac=round(gbs / pk["hbm"], 4),
                        us=round(dt * 1e6, 2), mode="exact (bit-identical to the reference's fp32 recurrence; default)"))
        if Ng <= 4096:      # the opt-in one-launch form: warp-level segmented scan + moments + normalisation
            st.gae_mode = "warp_scan"
            try:
                dt = time_kernel(lambda: st.compute_estimates(0.99, 0.95, True, True), iters=10)
            except Exception as e:          # an extra row must not take the headline line down
                out.append(dict(kernel="gae_scan_fused", error=repr(e)[:200]))
                del st
                continue
            gbs = 21 * T * Ng / dt / 1e9          # adv written once, never re-read: 21 B per (t, env)
            out.append(dict(kernel="gae_scan_fused (warp-level segmented scan + normalisation, one launch)", T=T,
                            n_envs=Ng, bytes_per_unit=21, bound="hbm", achieved=round(gbs, 1), peak=pk["hbm"],
                            unit="GB/s", frac=round(gbs / pk["hbm"], 4), us=round(dt * 1e6, 2),
                            mode="warp_scan (gae_mode='warp_scan'; within 1e-5 of the exact kernels)"))
        del st
    torch.cuda.empty_cache()
    return out


def gemm_roofline(agent, in_dim, hp, pk):
    """Roofline of the DOMINANT kernel of the PPO iteration (largest share of the timed region, profiles/breakdown):
    the update phase's tensor-core GEMM at the accumulation-window size.  One gather-sized forward + backward pass of
    the policy is run with every `tpp_gemm_tc` launch bracketed by CUDA events on the launching stream (the kernels
    are 50-250 us, the host stays ahead); `achieved` = sum of algorithmic FLOPs (2*M*N*K per launch; the 3xTF32 split
    passes are overhead, not work) / sum of durations over ALL launches of that kernel = its time-weighted average,
    against the measured dense bf16 peak.  `traffic` / `hbm_frac`: dram bytes per launch (average) from the committed
    `ncu --set full` capture of the same launches (profiles/ncu_traffic.json) over the live average duration."""
    from tpp_b200.common.engine import MLPEngineTC
    mb = min(hp["mini_batch_size"], hp["n_steps"] * hp["n_envs"] // hp["n_minibatch"])
    rows = mb * getattr(agent, "group_size", 1)   # rows of one launch: the minibatches of an accumulation window share a pass
    eng = agent.engine
    dev = agent.policy.flat.device
    raw = bool(getattr(eng, "raw_pixels", False))     # image observations reach layer 1 as integer pixel values
    if raw:
        x = torch.randint(0, 256, (rows, eng.ld_in), device=dev).float()
    else:
        x = torch.randn(rows, (in_dim + 3) // 4 * 4, device=dev)[:, :in_dim]
    dhead = torch.randn(rows, eng.ld_head, device=dev) / rows

    def fb():
        if raw:
            eng.forward(x, rows, raw=True)
        else:
            eng.forward(x, rows)
        eng.backward(dhead, rows)
    dt_all = time_kernel(fb, iters=10)
    agent.policy.flat_grad.zero_()
    macs = sum(l[2] * l[3] for l in eng.layers) + eng.latent * (eng.A + 1)
    flops_all = 6.0 * macs * rows - 2.0 * eng.layers[0][2] * eng.layers[0][3] * rows    # no dgrad for the first layer
    out = dict(bound="tensor", peak=pk["bf16"], unit="TFLOP/s", rows_per_launch=rows,
               policy_fwd_bwd_tflops=round(flops_all / dt_all / 1e12, 2), policy_fwd_bwd_us=round(dt_all * 1e6, 1),
               note="peak = dense bf16 cuBLAS (" + pk["source"] + "); a TF32 kernel tops out at 1/2 of it, the "
                    "fp32-parity 3xTF32 mode at 1/6; frac is the time-weighted average over every launch of the kernel in "
                    "one window pass, not the best launch")
    if not isinstance(eng, MLPEngineTC):
        out.update(kernel="gemm_f32_kernel (CUDA cores, exact fp32)", achieved=out["policy_fwd_bwd_tflops"],
                   frac=round(out["policy_fwd_bwd_tflops"] / pk["bf16"], 4), traffic=None)
        return out
    # per-launch events around every tensor-core GEMM of one pass
    records, inner = [], eng._tc

    def timed_tc(a, lda, b, ldb, M, N, K, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        inner(a, lda, b, ldb, M, N, K, **kw)
        e1.record()
        records.append((kw.get("block_n", 0), M, N, K, e0, e1))
    eng._tc = timed_tc
    try:
        for _ in range(4):
            fb()
        torch.cuda.synchronize()
        records.clear()
        reps = 5
        for _ in range(reps):
            fb()
        torch.cuda.synchronize()
    finally:
        del eng._tc
    agent.policy.flat_grad.zero_()
    by_tile = {}
    for bn, M, N, K, e0, e1 in records:
        bn = 513 if bn == 514 else bn      # the LEAN (3-stage, 8 epilogue warps) instantiation is the same kernel
        d = by_tile.setdefault(bn, dict(ms=0.0, flops=0.0, n=0, shapes=set()))
        d["ms"] += e0.elapsed_time(e1)
        d["flops"] += 2.0 * M * N * K
        d["n"] += 1
        d["shapes"].add((M, N, K))
    bn_dom = max(by_tile, key=lambda k: by_tile[k]["ms"])           # the tile variant with the largest time share
    d = by_tile[bn_dom]
    tf = d["flops"] / (d["ms"] * 1e-3) / 1e12
    tile = {0: "gemm_tc_kernel<128>", 64: "gemm_tc_kernel<64>", 128: "gemm_tc_kernel<128>", 256: "gemm_tc_kernel<256>",
            512: "gemm_tc_kernel<256, pair> (256x256 tile on a CTA pair, cta_group::2)",
            513: "gemm_tc_kernel<256, pair, persistent> (256x256 tiles on persistent CTA pairs, cta_group::2; forward + "
                 "weight-gradient launches in the 3-stage / 8-epilogue-warp instantiation)",
            65: "gemm_tc_kernel<64, pair, persistent>"}.get(bn_dom, f"gemm_tc_kernel<{bn_dom}>")
    us_avg = d["ms"] * 1e3 / d["n"]
    traffic = hbm_frac = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        per = [tj.get(f"gemm_tc_kernel[{M},{N},{K}]") for (M, N, K) in d["shapes"]]
        key = tj.get(f"window_pass[{rows}]")           # average dram bytes per launch of the dominant kernel in a pass
        if key is not None:
            traffic = key
        elif per and all(p is not None for p in per):
            traffic = float(np.mean(per))
        if traffic is not None:
            hbm_frac = round(traffic / (us_avg * 1e-6) / 1e9 / pk["hbm"], 4)
    share = d["ms"] / sum(v["ms"] for v in by_tile.values())
    out.update(kernel="%s (tcgen05 kind::tf32, %s)" % (tile, "3xTF32" if eng.precision == 3 else "1xTF32"),
               launches_per_pass=d["n"] // reps, shapes=sorted(d["shapes"]), achieved=round(tf, 2),
               frac=round(tf / pk["bf16"], 4), frac_of_sustained=round(tf / pk["bf16_sustained"], 4)
               if pk.get("bf16_sustained") else None, us_per_launch=round(us_avg, 2), traffic=traffic,
               hbm_frac=hbm_frac, share_of_gemm_time_in_pass=round(share, 3))
    return out


def measure(name, hp, rank, local, world, args, logger=True, extra=None):
    """Device-timed (`ms`) and end-to-end (`ms_e2e`) PPO iterations of one workload on this rank's shard."""
    from tpp_b200.common.logger import Logger
    device = f"cuda:{local}"
    agent, in_dim = build_agent(name, hp, rank, device, matmul=args.matmul,
                                fuse_accum=args.fuse_accum if args.fuse_accum == "auto" else int(args.fuse_accum),
                                rollout_chains=args.rollout_chains, **(extra or {}))
    if world > 1:
        agent.shard(world)
    N, T = hp["n_envs"], hp["n_steps"]
    st, env = agent.storage, agent.env

    def counters():
        return agent.n_launches + agent.engine.n_launches + st.n_launches + agent.optimizer.n_launches

    def iteration():
        agent.collect_rollout(env, st)
        st.compute_estimates(agent.gamma, agent.lmbda, True, True)
        agent.optimize()
        agent._carry_over(st)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing: minibatch permutations pre-generated and pre-uploaded -------------------
    env.reset_rollout(st)
    mb = min(hp["mini_batch_size"], T * N // hp["n_minibatch"])
    pre = [torch.randperm(T * N)[:(T * N) // mb * mb].view(-1, mb).to(device) for _ in range(hp["epoch"])]
    cyc = {"i": 0}

    def resident_indices(_mb):
        cyc["i"] += 1
        return pre[cyc["i"] % len(pre)]
    st.epoch_indices = resident_indices
    for _ in range(max(args.warmup, 3)):
        iteration()
    barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    l0 = counters()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.profiler.start()          # `ncu --profile-from-start off` captures exactly the timed region
    e0.record()
    for _ in range(args.steps):
        iteration()
    e1.record()
    barrier()
    torch.cuda.profiler.stop()
    ms = e0.elapsed_time(e1)
    launches = counters() - l0
    clk = clocks.stop() if rank == 0 else None
    res = dict(agent=agent, in_dim=in_dim, ms=ms, launches=launches, clocks=clk, mb=mb)
    if args.timed_region_only:
        return res
    # phase split (CUDA events, outside the timed region): the rollout alone
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    r0.record()
    for _ in range(3):
        agent.collect_rollout(env, st)
    r1.record()
    torch.cuda.synchronize()
    res["rollout_ms"] = r0.elapsed_time(r1) / 3
    # ---- end to end through the public API: PPO.train() with host-side index generation + upload + readbacks -
    del st.__dict__["epoch_indices"]
    if logger:
        agent.logger = Logger(N)                  # this package's Logger: device-side episode accounting
        agent.logger.max_steps = hp.get("max_steps", 500)
    agent.t = 0
    agent.train(T * N * 2)                              # warm the API path (graph already captured)
    barrier()
    agent.t = 0
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter()
    t0.record()
    agent.train(T * N * args.steps)
    t1.record()
    barrier()
    wall = time.perf_counter() - w0
    ms_e2e = max(t0.elapsed_time(t1), wall * 1e3)
    times = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=device)
    in_sync = None
    if world > 1:
        torch.distributed.all_reduce(times, op=torch.distributed.ReduceOp.MAX)
        # every rank applied the same averaged gradient: the parameter vectors must be bit-identical
        chk = agent.policy.flat.double().sum().reshape(1)
        lo, hi = chk.clone(), chk.clone()
        torch.distributed.all_reduce(lo, op=torch.distributed.ReduceOp.MIN)
        torch.distributed.all_reduce(hi, op=torch.distributed.ReduceOp.MAX)
        in_sync = bool((lo == hi).item())
    res["ms"], res["ms_e2e"] = times.tolist()
    n_mb = (T * N) // mb
    res.update(in_sync=in_sync, h2d=hp["epoch"] * n_mb * mb * 4 + 8 + 32,       # int32 indices + lr + loss coefficients
               d2h=hp["epoch"] * n_mb * 20 * 8 + (2 + 80) * 8,                 # loss statistics + episode records
               episodes=getattr(agent.logger, "num_episodes", None))
    return res


def run_ours(args):
    from tpp_b200 import parallel
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local)
    device = f"cuda:{local}"
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=torch.device(device))
    name = args.workload
    hp_total = workload_hp(name)
    strong = name == "boxworld"        # configs[1] names a TOTAL env count: shard it; cartpole stays per-GPU (weak)
    hp = parallel.shard_hyperparameters(hp_total, world) if (strong and world > 1) else hp_total
    res = measure(name, hp, rank, local, world, args)
    if args.timed_region_only:
        if rank == 0:
            print(json.dumps({"ms_per_step": res["ms"] / args.steps, "gpu_launches": int(res["launches"])}))
        return
    weak = None
    if strong and world > 1:           # the same run, 4096 envs PER GPU (round 1's headline), as an extra key
        del res["agent"]
        torch.cuda.empty_cache()
        w = measure(name, hp_total, rank, local, world, args)
        tot = hp_total["n_envs"] * hp_total["n_steps"] * args.steps * world
        weak = {"scaling": "weak", "n_envs_per_gpu": hp_total["n_envs"], "value": round(tot / (w["ms"] * 1e-3), 1),
                "ms_per_step": round(w["ms"] / args.steps, 3), "e2e_value": round(tot / (w["ms_e2e"] * 1e-3), 1),
                "replicas_in_sync": w["in_sync"]}
        res_agent = w["agent"]
    else:
        res_agent = res["agent"]
    if rank != 0:
        torch.distributed.barrier()          # rank 0 is done with its single-GPU legs (roofline, CPU baseline)
        finish(world)
        return
    pk = peaks()
    N, T = hp["n_envs"], hp["n_steps"]
    total_steps = N * T * args.steps * world
    value = total_steps / (res["ms"] * 1e-3)
    e2e_value = total_steps / (res["ms_e2e"] * 1e-3)
    roof = gemm_roofline(res_agent, res["in_dim"], hp_total if weak else hp, pk)
    extra = kernel_rooflines(pk, device) if (not args.no_kernel_rooflines and world == 1) else []
    cpu = cpu_baseline(name, budget_s=20.0) if not args.no_cpu_baseline else None
    configs = None
    if world == 1 and name == "boxworld" and not args.no_configs:
        configs = other_configs(args, rank, local, pk)
    n_e, n_s = CPU_SAMPLE[name]
    line = {
        "metric": METRIC, "value": round(value, 1), "unit": "env-steps/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": round(res["ms"] / args.steps, 3),
        "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": workload_string(name, hp_total),
                   "parallelism": f"env-sharded dp{world}: {N} envs and minibatch {res['mb']} per GPU "
                                  f"(x{getattr(res_agent, 'group_size', 1)} minibatches per pass: gradient-accumulation "
                                  f"window), {args.matmul}",
                   "yaml_set": {"boxworld": "boxworld-impala", "cartpole": "cartpole"}[name],
                   "l2": "rollout + minibatch working set > L2 (inputs larger than 126 MB)" if name == "boxworld"
                   else "small working set (latency-bound)",
                   "reference_arm": f"--impl reference times the CPU port on the same workload with n_steps={n_s} of "
                                    f"{hp_total['n_steps']} per CPU step (n_envs={n_e}, minibatch and epochs unchanged; "
                                    "every phase is linear in n_steps, so env-steps/s is the same quantity)"},
        "e2e": {"value": round(e2e_value, 1), "unit": "env-steps/s", "h2d_bytes_per_step": res["h2d"],
                "d2h_bytes_per_step": res["d2h"], "ms_per_step": round(res["ms_e2e"] / args.steps, 3),
                "episodes_logged": res["episodes"],
                "api": "PPO.train() with tpp_b200 Logger: host randperm -> pinned int32 H2D per epoch (copy stream, "
                       "overlapping the previous epoch); D2H loss statistics + device-side episode records (last 40 "
                       "episodes + count) into pinned buffers, read by the host after the next rollout is enqueued"},
        "phases": {"rollout_ms": round(res["rollout_ms"], 3),
                   "us_per_rollout_step": round(res["rollout_ms"] * 1e3 / T, 2),
                   "gae_plus_update_ms": round(res["ms"] / args.steps - res["rollout_ms"], 3)},
        "gpu_launches": int(res["launches"]), "clocks": res["clocks"], "replicas_in_sync": res["in_sync"],
        "roofline": roof, "weak_scaling": weak, "kernel_rooflines": extra, "cpu_baseline": cpu, "configs": configs,
    }
    print(json.dumps(line))
    if world > 1:
        torch.distributed.barrier()
    finish(world)


def other_configs(args, rank, local, pk):
    """BASELINE configs[0] (C1) and configs[3] (C4) measured in the same run (N=1)."""
    out = {}
    a = argparse.Namespace(**vars(args))
    a.steps, a.warmup = max(args.steps, 10), 3
    hp = workload_hp("cartpole")
    r = measure("cartpole", hp, rank, local, 1, a)
    tot = hp["n_envs"] * hp["n_steps"] * a.steps
    c1 = {"workload": workload_string("cartpole", hp), "yaml_set": "cartpole",
          "value": round(tot / (r["ms"] * 1e-3), 1), "ms_per_step": round(r["ms"] / a.steps, 3),
          "e2e": {"value": round(tot / (r["ms_e2e"] * 1e-3), 1), "ms_per_step": round(r["ms_e2e"] / a.steps, 3),
                  "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": r["d2h"]},
          "gpu_launches": int(r["launches"]), "unit": "env-steps/s", "steps": a.steps,
          "rollout_ms": round(r["rollout_ms"], 3), "us_per_rollout_step": round(r["rollout_ms"] * 1e3 / hp["n_steps"], 2)}
    del r
    torch.cuda.empty_cache()
    if not args.no_cpu_baseline:
        c1["cpu_baseline"] = cpu_baseline("cartpole", budget_s=12.0)       # the SAME 256 x 256 on the host cores
    out["C1_cartpole"] = c1
    try:
        p = argparse.Namespace(**vars(args))
        p.steps, p.warmup = 2, 1
        out["C4_procgen"] = run_procgen(p, emit=False)
    except Exception as e:      # the headline line must survive a failure of an extra config
        out["C4_procgen"] = {"error": repr(e)[:300]}
    torch.cuda.empty_cache()
    try:                        # configs[4]: the hard-500 set as it is (256 envs, minibatch 8192), throughput only
        p = argparse.Namespace(**vars(args))
        p.steps, p.warmup = 2, 1
        out["C5_procgen_hard500"] = run_procgen(p, emit=False, name="procgen_c5", with_roofline=False)
    except Exception as e:
        out["C5_procgen_hard500"] = {"error": repr(e)[:300]}
    torch.cuda.empty_cache()
    return out


class SyntheticProcgen:
    """Host-stepped VecEnv with Procgen's contract after VecExtractDictObs (uint8 64x64x3 frames, 15 actions, float
    rewards): cycles a pool of pre-generated frames, Bernoulli(0.01)*10 rewards, Bernoulli(1/200) dones at
    no CPU cost.  Wrapped in tpp_b200's StagedVecEnv (device VecNormalize, frames staged as uint8)."""

    def __init__(self, n, pool=64, seed=0):
        from tpp_b200.discrete_env.pre_vec_env import Box, Discrete
        rng = np.random.default_rng(seed)
        self.num_envs = n
        self.frames = rng.integers(0, 256, (pool, n, 64, 64, 3), dtype=np.uint8)
        self.rew = ((rng.random((pool, n)) < 0.01) * 10.0).astype(np.float32)
        self.done = rng.random((pool, n)) < (1 / 200)
        self.i = 0
        self.observation_space = Box(np.zeros((64, 64, 3)), np.full((64, 64, 3), 255), dtype=np.uint8)
        self.action_space = Discrete(15)

    def reset(self):
        return self.frames[0]

    def step(self, act):
        self.i = (self.i + 1) % len(self.frames)
        return self.frames[self.i], self.rew[self.i], self.done[self.i], None

    def close(self):
        pass


def run_procgen(args, emit=True, name="procgen", with_roofline=True):
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local)
    device = f"cuda:{local}"
    if world > 1 and not torch.distributed.is_initialized():
        torch.distributed.init_process_group("nccl", device_id=torch.device(device))
    from tpp_b200 import _lib
    from tpp_b200.agents.ppo import PPO
    from tpp_b200.common.env.procgen_wrappers import StagedVecEnv
    from tpp_b200.common.logger import Logger
    from tpp_b200.common.model import ImpalaModel
    from tpp_b200.common.policy import CategoricalPolicy
    from tpp_b200.common.storage import Storage
    hp = workload_hp(name)
    N, T = hp["n_envs"], hp["n_steps"]
    matmul = args.matmul if args.matmul in ("tf32x3", "tf32") else "tf32x3"
    env = StagedVecEnv(SyntheticProcgen(N, seed=rank), normalize_rew=hp.get("normalize_rew", True), gamma=hp["gamma"],
                       device=device)
    torch.manual_seed(6033)
    pol = CategoricalPolicy(ImpalaModel(3), False, 15).to(device).flatten_()
    st = Storage((3, 64, 64), 256, T, N, device)
    lg = Logger(N)
    lg.max_steps = 1000
    agent = PPO(env, pol, lg, st, device, 0, **{k: hp[k] for k in PPO_KEYS}, sample_seed=17, matmul=matmul)
    if world > 1:
        agent.shard(world)
    A = agent.n_actions

    def counters():
        return agent.n_launches + agent.engine.n_launches + st.n_launches + agent.optimizer.n_launches

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def resident_iteration():          # frames of the last rollout are in st.frames: policy rollout + GAE + update
        agent.engine.refresh_weights()
        for t in range(T):
            agent._host_step_device(st, t, N)
        head = agent._policy_head(st.obs_slot(T), st)
        st.value[T, :N] = head[:N, A]
        _lib.call("tpp_tick_advance", _lib.ptr(agent._tick), T, _lib.stream_ptr())
        st.compute_estimates(agent.gamma, agent.lmbda, True, True)
        agent.optimize()

    agent.train(T * N * 2)             # eager iteration, then the iteration that captures the graphs
    for _ in range(max(args.warmup, 1)):
        resident_iteration()
    barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    l0 = counters()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        resident_iteration()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = counters() - l0
    clk = clocks.stop() if rank == 0 else None
    st.h2d_bytes, st.d2h_bytes = 0, 0
    agent.t = 0
    w0 = time.perf_counter()
    agent.train(T * N * args.steps)
    barrier()
    ms_e2e = (time.perf_counter() - w0) * 1e3
    h2d = st.h2d_bytes // args.steps
    d2h = st.d2h_bytes // args.steps + 24 * 20 * 8 + 82 * 8
    times = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=device)
    if world > 1:
        torch.distributed.all_reduce(times, op=torch.distributed.ReduceOp.MAX)
    ms, ms_e2e = times.tolist()
    if rank != 0:
        if world > 1 and emit:
            torch.distributed.barrier()
            finish(world)
        return None
    total = N * T * args.steps * world
    if not with_roofline:           # configs-block entry: the throughput numbers only (C4's entry carries the roofline)
        return {"workload": workload_string(name, hp), "yaml_set": "hard-500", "value": round(total / (ms * 1e-3), 1),
                "unit": "env-steps/s", "ms_per_step": round(ms / args.steps, 3), "steps": args.steps,
                "e2e": {"value": round(total / (ms_e2e * 1e-3), 1), "ms_per_step": round(ms_e2e / args.steps, 3),
                        "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
                "gpu_launches": int(launches), "clocks": clk}
    pk = peaks()
    sys.path.insert(0, os.path.join(ROOT, "profiles"))
    import run_conv_kernels as rck
    forms = rck.time_forms()
    pixels = rck.SHAPE["B"] * rck.SHAPE["H"] * rck.SHAPE["W"]
    flops = 2.0 * pixels * 9 * rck.SHAPE["C"] ** 2
    tf = flops / forms["forward"] / 1e6
    traffic = None
    tj = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tj):
        traffic = json.load(open(tj)).get("gemm_tc_conv_forward")
    roof = {"kernel": "gemm_tc_kernel<16>, implicit 3x3 convolution forward (TMA im2col, 3xTF32), IMPALA block 1: "
                      "2048 x 32x32 pixels, 16 -> 16 channels", "bound": "tensor", "achieved": round(tf, 2),
            "peak": pk["bf16"], "unit": "TFLOP/s", "frac": round(tf / pk["bf16"], 4), "traffic": traffic,
            "us_per_launch": round(forms["forward"], 1), "dgrad_us": round(forms["dgrad"], 1),
            "wgrad_us": round(forms["wgrad_fma"], 1), "wgrad_tensor_core_form_us": round(forms["wgrad"], 1),
            "wgrad_note": "weight gradients of the narrow layers run on the fp32 FMA pipe (csrc/conv_cc.cu, exact fp32, "
                          "every operand byte staged once): 2*pixels*9*Cin*Cout FLOP at "
                          f"{flops / forms['wgrad_fma'] / 1e6:.1f} TFLOP/s; the tcgen05 im2col form of the same "
                          "contraction is kept for other shapes",
            "gathered_operand_GBps": round(18 * pixels * rck.SHAPE["C"] * 4 / forms["forward"] / 1e3, 1),
            "note": f"peak = dense bf16 ({pk['source']}); algorithmic FLOPs 2*pixels*9*Cin*Cout; the tile is bound by "
                    "the fixed cost of its 54 MMAs of 128 x 16 x 8 per tile (an MMA with N = 16 takes as long as a wide one); "
                    "staging the input rows once per filter column -- a third of the L2 gather -- left the time "
                    "unchanged (profiles/conv_halo_experiment_r02.patch: 345 vs 340 us)"}
    cpu = cpu_baseline(name, budget_s=15.0) if not args.no_cpu_baseline else None
    n_e, n_s = CPU_SAMPLE[name]
    line = {
        "metric": METRIC, "value": round(total / (ms * 1e-3), 1),
        "unit": "env-steps/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 1),
        "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_string(name, hp), "yaml_set": "hard-500",
                   "parallelism": f"env-sharded dp{world}: {N} envs per GPU, {matmul}",
                   "l2": "minibatch activations (GBs) >> L2",
                   "reference_arm": f"CPU port on the same workload with n_steps={n_s} of {T} per CPU step (n_envs={n_e})"},
        "e2e": {"value": round(total / (ms_e2e * 1e-3), 1), "unit": "env-steps/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "ms_per_step": round(ms_e2e / args.steps, 3),
                "api": "PPO.train() with a host-stepped env behind StagedVecEnv: per step double-buffered pinned uint8 "
                       "frames H2D, actions D2H (event wait); rewards / dones one H2D per rollout; VecNormalize on the "
                       "device; tpp_b200 Logger"},
        "gpu_launches": int(launches), "clocks": clk, "roofline": roof, "cpu_baseline": cpu}
    if emit:
        print(json.dumps(line))
        if world > 1:
            torch.distributed.barrier()
        finish(world)
    return line


# --------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's numpy/torch path (the reference itself is Python and cannot
# travel to the GPU box; oracle/* restates it and is pinned against it bit-for-bit, see tests/)
# --------------------------------------------------------------------------------------------------

def _cpu_setup(name, n_envs, n_steps):
    from oracle import boxworld as obw
    from oracle import ppo as oppo
    from oracle.prevec import OraclePreVec
    hp = dict(workload_hp(name), n_envs=n_envs, n_steps=n_steps)
    torch.manual_seed(6033)
    kw = dict(epoch=hp["epoch"], n_minibatch=hp["n_minibatch"], mini_batch_size=hp["mini_batch_size"],
              grad_clip_norm=hp["grad_clip_norm"], eps_clip=hp["eps_clip"], value_coef=hp["value_coef"],
              entropy_coef=hp["entropy_coef"])
    if name.startswith("procgen"):
        rng = np.random.default_rng(0)
        pool = rng.integers(0, 256, (8, n_envs, 64, 64, 3), dtype=np.uint8)
        state = {"i": 0}

        def env_step(a):
            state["i"] = (state["i"] + 1) % len(pool)
            return pool[state["i"]], (rng.random(n_envs) < 0.01) * 10.0, rng.random(n_envs) < (1 / 200)
        tf = lambda f: np.ascontiguousarray(f.transpose(0, 3, 1, 2)).astype(np.float32) / 255.0
        pol = oppo.OraclePolicy(oppo.OracleImpala(3), 15)
        opt = oppo.make_adam(pol, hp["learning_rate"])

        def iteration(obs):
            obs, _ = oppo.ppo_iteration(env_step, obs, pol, opt, n_steps, n_envs, hp["gamma"], hp["lmbda"],
                                        obs_transform=tf, **kw)
            return obs
        return iteration, pool[0]
    if name == "boxworld":
        env = obw.BoxWorldOracle(n_envs, hp["grid_size"], hp["goal_length"], hp["num_distractor"],
                                 hp["distractor_length"], max_steps=hp["max_steps"], start_seed=6033, n_levels=500)
        vn = obw.VecNormalizeOracle(n_envs)

        def env_step(a):
            w, r, d = env.step(a)
            return w, vn.step(r.astype(np.float64), d), d
        tf = lambda w: obw.frame_to_obs(w).reshape(n_envs, -1)
        obs0, in_dim, A = env.world, 3 * (hp["grid_size"] + 2) ** 2, 4
    else:
        env = OraclePreVec("cartpole", n_envs, seed=6033)
        env.reset()

        def env_step(a):
            return env.step(a)
        tf, obs0, in_dim, A = None, env.obs(), 9, 2
    pol = oppo.OraclePolicy(oppo.OracleMLP(in_dim, hp["depth"], hp["mid_weight"], hp["latent_size"]), A)
    opt = oppo.make_adam(pol, hp["learning_rate"])

    def iteration(obs):
        obs, _ = oppo.ppo_iteration(env_step, obs, pol, opt, n_steps, n_envs, hp["gamma"], hp["lmbda"],
                                    obs_transform=tf, **kw)
        return obs
    return iteration, obs0


def _host_threads():
    """torchrun exports OMP_NUM_THREADS=1: give the CPU arm every host core it can use."""
    n = os.cpu_count() or 1
    if torch.get_num_threads() < n:
        torch.set_num_threads(n)
    return torch.get_num_threads()


def cpu_baseline(name, budget_s=20.0, steps=None):
    """Time the oracle port on the host cores on a bounded sample of the same workload."""
    threads = _host_threads()
    n_envs, n_steps = CPU_SAMPLE[name]
    full = workload_hp(name)
    iteration, obs = _cpu_setup(name, n_envs, n_steps)
    obs = iteration(obs)                     # warm-up
    t0, k = time.perf_counter(), 0
    while True:
        obs = iteration(obs)
        k += 1
        el = time.perf_counter() - t0
        if (steps is not None and k >= steps) or (steps is None and (el > budget_s or k >= 50)):
            break
    same = (n_envs, n_steps) == (full["n_envs"], full["n_steps"])
    return {"value": round(n_envs * n_steps * k / el, 1), "unit": "env-steps/s", "cores": threads,
            "host_cpus": os.cpu_count(), "kind": "port", "same_config": same,
            "sample": f"{k} PPO iterations of the oracle port at n_envs={n_envs}, n_steps={n_steps}"
                      + ("" if same else f" (of {full['n_steps']})")
                      + f" (same YAML hyperparameters: minibatch, epochs); numpy env step is single-threaded, torch uses "
                      f"{threads} threads"}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    threads = _host_threads()
    name = args.workload
    n_envs, n_steps = CPU_SAMPLE[name]
    full = workload_hp(name)
    iteration, obs = _cpu_setup(name, n_envs, n_steps)
    for _ in range(max(1, min(args.warmup, 2))):
        obs = iteration(obs)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        obs = iteration(obs)
    el = time.perf_counter() - t0
    value = n_envs * n_steps * args.steps / el
    sample = (f"{args.steps} PPO iterations at n_envs={n_envs}, n_steps={n_steps} of {full['n_steps']} on the host CPU "
              f"({threads} torch threads; numpy env step single-threaded)")
    cpu = {"value": round(value, 1), "unit": "env-steps/s", "cores": threads, "kind": "port", "sample": sample}
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": round(value, 1),
        "unit": "env-steps/s", "n_gpus": int(os.environ.get("WORLD_SIZE", 1)), "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(el / args.steps * 1e3, 3), "higher_is_better": True,
        "scaling": "strong" if name == "boxworld" else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_string(name, full),
                   "yaml_set": {"boxworld": "boxworld-impala", "cartpole": "cartpole", "procgen": "hard-500",
                                "procgen_c5": "hard-500"}[name],
                   "parallelism": f"host CPU, {threads} threads (not sharded: rank 0 only)",
                   "reference_arm": f"CPU port of the reference path on the same workload with n_steps={n_steps} of "
                                    f"{full['n_steps']} per CPU step (n_envs={n_envs}, minibatch and epochs unchanged; "
                                    "every phase is linear in n_steps, so env-steps/s is the same quantity)"},
        "cpu_baseline": cpu,
        "e2e": {"value": round(value, 1), "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def finish(world):
    """Leave without tearing the NCCL communicator down: CUDA graphs that captured its all-reduce are still alive, and
    destroy_process_group() then waits forever (observed on 2 GPUs).  Every rank has passed the last barrier."""
    sys.stdout.flush()
    sys.stderr.flush()
    if world > 1:
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="boxworld", choices=["boxworld", "cartpole", "procgen", "procgen_c5"])
    ap.add_argument("--matmul", default="tf32x3", choices=["tf32x3", "tf32", "fp32"],
                    help="dense-layer arithmetic: tcgen05 3xTF32 (fp32-parity, default), tcgen05 single TF32, CUDA-core fp32")
    ap.add_argument("--fuse-accum", default="auto",
                    help="minibatches of one gradient-accumulation window sharing a forward/backward pass (auto = all)")
    ap.add_argument("--rollout-chains", type=int, default=1,
                    help="env ranges stepping through the rollout as concurrent kernel chains (experiment: no gain, "
                         "the T sequential steps are latency-bound whatever the range size)")
    ap.add_argument("--no-kernel-rooflines", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU leg (profiling runs only)")
    ap.add_argument("--no-configs", action="store_true", help="skip the C1 / C4 block of the default run")
    ap.add_argument("--timed-region-only", action="store_true", help="stop after the device-timed loop (ncu runs)")
    ap.add_argument("--only-kernel-rooflines", action="store_true", help="time just the HBM-bound kernels")
    args = ap.parse_args()
    if args.only_kernel_rooflines:
        print(json.dumps(kernel_rooflines(peaks(), "cuda:0")))
    elif args.impl == "reference":
        run_reference(args)
    elif args.workload.startswith("procgen"):
        run_procgen(args, name=args.workload)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
