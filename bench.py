#!/usr/bin/env python
"""Benchmark of the PPO rollout-and-update hot path (BASELINE.json metric: PPO env-steps/sec, rollout+GAE+update).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload boxworld|cartpole|procgen]

One "step" = one full PPO iteration: T fused rollout steps over the rank's envs (policy forward -> Philox action
sampling -> env step kernel writing into the rollout), bootstrap value, GAE + advantage normalisation, and
`epoch` x minibatch updates (device gather -> policy fwd -> fused loss fwd+bwd -> policy bwd -> clip+Adam).

Default workload = BASELINE.json configs[1]: boxworld_env_vec PPO, n_envs=4096 per GPU (weak scaling), grid 12,
goal 5, 3x3 distractors, 500-level bank, T=256, 3 epochs, n_minibatch 8, mini_batch_size 8192 (the reference's
`boxworld-impala` YAML set, hyperparams/procgen/config.yml:575-601), with an MLP policy on the flattened
3x14x14 frame (the reference has no working Box-World policy; choice documented in DESIGN.md).

`--workload procgen` = BASELINE configs[3] shape (coinrun hard-500: IMPALA-CNN, 64 envs per GPU, 64x64x3 uint8
frames) with a synthetic host engine (the Procgen engine is not in this image): `value` = policy rollout on frames
resident in HBM + GAE + update, `e2e` = `PPO.train()` with every step's frames staged from pinned host memory and the
actions read back.

Prints ONE JSON line (rank 0).  `value` = device-timed iterations with inputs resident (minibatch permutations
pre-uploaded); `e2e` = the same iterations through the public API `PPO.train()` with per-epoch index upload from
pinned host memory and device->host reads of the loss summary and the logger's reward/done batches.
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

WORKLOADS = {
    # reference YAML set `boxworld-impala` (config.yml:575-601) at BASELINE configs[1]'s n_envs
    "boxworld": dict(n_envs=4096, n_steps=256, epoch=3, n_minibatch=8, mini_batch_size=8192, gamma=0.999, lmbda=0.95,
                     learning_rate=5e-4, grad_clip_norm=0.5, eps_clip=0.2, value_coef=0.5, entropy_coef=0.01,
                     grid_size=12, goal_length=5, num_distractor=3, distractor_length=3, max_steps=1000,
                     normalize_rew=True, depth=4, mid_weight=256, latent_size=64),
    # reference YAML set `cartpole` (config.yml:972-995): BASELINE configs[0], the reference's CPU-runnable case
    "cartpole": dict(n_envs=256, n_steps=256, epoch=3, n_minibatch=16, mini_batch_size=8192, gamma=0.99, lmbda=0.95,
                     learning_rate=5e-4, grad_clip_norm=0.5, eps_clip=0.2, value_coef=0.5, entropy_coef=0.02,
                     depth=4, mid_weight=256, latent_size=64),
    # reference YAML set `hard-500` (config.yml:81-99) as launched for coinrun / maze / heist: BASELINE configs[3]
    "procgen": dict(n_envs=64, n_steps=256, epoch=3, n_minibatch=8, mini_batch_size=8192, gamma=0.999, lmbda=0.95,
                    learning_rate=5e-4, grad_clip_norm=0.5, eps_clip=0.2, value_coef=0.5, entropy_coef=0.01),
}
PPO_KEYS = ("n_steps", "n_envs", "epoch", "n_minibatch", "mini_batch_size", "gamma", "lmbda", "learning_rate",
            "grad_clip_norm", "eps_clip", "value_coef", "entropy_coef")


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

def build_agent(name, hp, rank, device):
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
        obs_shape = env.observation_space.shape
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
    agent = PPO(env, pol, None, st, device, 0, **{k: hp[k] for k in PPO_KEYS}, sample_seed=17 + rank,
                matmul=hp.get("matmul", "tf32x3"), fuse_accum=hp.get("fuse_accum", "auto"),
                rollout_chains=hp.get("rollout_chains", 1))
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
            out.append(dict(kernel=f"env_step_{env.family}", n_envs=N, bytes_per_unit=bytes_per_step, bound="hbm",
                            achieved=round(gbs, 1), peak=pk["hbm"], unit="GB/s", frac=round(gbs / pk["hbm"], 4),
                            env_steps_per_s=round(N / dt, 1), us=round(dt * 1e6, 2),
                            working_set_mb=round(2 * env.n_obs * 4 * N / 1e6, 1)))
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
                        us=round(dt * 1e6, 2)))
        del st
    torch.cuda.empty_cache()
    return out


def gemm_roofline(agent, in_dim, hp, pk):
    """Roofline of the dominant kernel of the PPO iteration: the policy's dense-layer GEMM.  Timed live with CUDA
    events on the launching stream: the largest layer of the update (layer 1 forward, M = minibatch, N = 256,
    K = in_dim) launched alone, algorithmic FLOPs = 2*M*N*K per launch (the 3xTF32 split passes are overhead, not
    algorithmic work), against the measured dense bf16 tensor peak.  Also reports the whole policy forward+backward
    (all its launches) in TFLOP/s."""
    from tpp_b200.common.engine import MLPEngineTC
    from tpp_b200._lib import EPI_BIAS, EPI_RELU, TC_A_EXACT, ptr
    mb = min(hp["mini_batch_size"], hp["n_steps"] * hp["n_envs"] // hp["n_minibatch"])
    mb *= getattr(agent, "group_size", 1)     # rows of one launch: the minibatches of an accumulation window share a pass
    eng = agent.engine
    ld = (in_dim + 3) // 4 * 4
    raw = bool(getattr(eng, "raw_pixels", False))     # image observations reach layer 1 as integer pixel values
    if raw:
        x = torch.randint(0, 256, (mb, eng.ld_in), device=agent.policy.flat.device).float()
    else:
        x = torch.randn(mb, ld, device=agent.policy.flat.device)[:, :in_dim]
    dhead = torch.randn(mb, eng.ld_head, device=x.device) / mb

    def fb():
        if raw:
            eng.forward(x, mb, raw=True)
        else:
            eng.forward(x, mb)
        eng.backward(dhead, mb)
    dt_all = time_kernel(fb, iters=10)
    agent.policy.flat_grad.zero_()
    macs = sum(l[2] * l[3] for l in eng.layers) + eng.latent * (eng.A + 1)
    flops_all = 6.0 * macs * mb - 2.0 * eng.layers[0][2] * eng.layers[0][3] * mb    # no dgrad for the first layer
    w_off, b_off, fin, fout, relu = eng.layers[0]
    ws = eng._workspace(mb)
    if isinstance(eng, MLPEngineTC):
        w, h = (eng.w0_raw if raw else eng.w[0]), ws.h[0]
        a = (x, x) if raw else (ws.x["hi"], ws.x["lo"])
        bn = eng._bn(mb, fout)
        one = lambda: eng._tc(a, eng.ld_in, (w["hi"], w["lo"]), w["ldk"], mb, fout, fin,
                              flags=EPI_BIAS | EPI_RELU, bias=eng._p(b_off), out_pair=(h["hi"], h["lo"]), ldc=h["ld"],
                              exact=TC_A_EXACT if raw else 0, block_n=bn)
        tile = {0: "gemm_tc_kernel<128>", 64: "gemm_tc_kernel<64>", 256: "gemm_tc_kernel<256>",
                512: "gemm_tc_kernel<256, pair> (256x256 tile on a CTA pair, cta_group::2)",
                513: "gemm_tc_kernel<256, pair, persistent> (256x256 tiles on persistent CTA pairs, cta_group::2)"}[bn]
        name = "%s (tcgen05 kind::tf32, %s%s)" % (
            tile, "3xTF32" if eng.precision == 3 else "1xTF32",
            ", pixel operand exact in TF32: 2 of the 3 passes" if raw and eng.precision == 3 else "")
    else:
        one = lambda: eng._gemm(ptr(x), x.stride(0), 1, eng._p(w_off), fin, 1, ptr(ws.acts[0]), fout, eng._p(b_off),
                                None, mb, fout, fin, EPI_BIAS | EPI_RELU)
        name = "gemm_f32_kernel (CUDA cores, exact fp32)"
    dt = time_kernel(one, iters=50)
    tf = 2.0 * mb * fout * fin / dt / 1e12
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):    # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu capture
        traffic = json.load(open(tpath)).get(f"gemm_tc_kernel[{mb},{fout},{fin}]")
    return dict(kernel=name, shape=[mb, fout, fin], bound="tensor", achieved=round(tf, 2), peak=pk["bf16"],
                unit="TFLOP/s", frac=round(tf / pk["bf16"], 4), traffic=traffic, us_per_launch=round(dt * 1e6, 2),
                policy_fwd_bwd_tflops=round(flops_all / dt_all / 1e12, 2),
                policy_fwd_bwd_us=round(dt_all * 1e6, 1),
                note="peak = dense bf16 cuBLAS (" + pk["source"] + "); a TF32 kernel tops out at 1/2 of it, the "
                     "fp32-parity 3xTF32 mode at 1/6")


def run_ours(args):
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local)
    device = f"cuda:{local}"
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=torch.device(device))
    hp = dict(WORKLOADS[args.workload], matmul=args.matmul,
              fuse_accum=args.fuse_accum if args.fuse_accum == "auto" else int(args.fuse_accum),
              rollout_chains=args.rollout_chains)
    agent, in_dim = build_agent(args.workload, hp, rank, device)
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
    n_perm = hp["epoch"] * (args.warmup + args.steps + 2)
    pre = [torch.randperm(T * N)[:(T * N) // mb * mb].view(-1, mb).to(device) for _ in range(hp["epoch"])]
    real_epoch_indices = st.epoch_indices
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

    if args.timed_region_only:
        if rank == 0:
            print(json.dumps({"ms_per_step": ms / args.steps, "gpu_launches": int(launches)}))
        return
    # ---- end to end through the public API: PPO.train() with host-side index generation + upload + readbacks -
    st.epoch_indices = real_epoch_indices

    class NullLogger:      # the reference's Logger consumes (rew_batch, done_batch) on the host every iteration
        logdir = None
        episode_reward_buffer = [0.0]

        def feed(self, *a):
            self.n = sum(x.size for x in a[:2])

        def dump(self, *a):
            pass
    agent.logger = NullLogger()
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
    ms, ms_e2e = times.tolist()

    if rank != 0:
        if world > 1:
            torch.distributed.destroy_process_group()
        return
    pk = peaks()
    total_steps = N * T * args.steps * world
    value = total_steps / (ms * 1e-3)
    e2e_value = total_steps / (ms_e2e * 1e-3)
    n_mb = (T * N) // mb
    h2d = hp["epoch"] * n_mb * mb * 8 + 8
    d2h = hp["epoch"] * n_mb * 20 * 8 + T * st.ld * (4 + 1)      # loss statistics + (raw reward f32, done u8) batches
    roof = gemm_roofline(agent, in_dim, hp, pk)
    extra = kernel_rooflines(pk, device) if not args.no_kernel_rooflines else []
    cpu = cpu_baseline(args.workload, budget_s=20.0) if not args.no_cpu_baseline else None
    line = {
        "metric": "PPO env-steps/sec (rollout+GAE+update)", "value": round(value, 1), "unit": "env-steps/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(ms / args.steps, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}_env_vec PPO: n_envs={N}/GPU, n_steps={T}, "
                               f"epoch={hp['epoch']}, minibatch={mb} (x{getattr(agent, 'group_size', 1)} per pass: "
                               f"gradient-accumulation window), MLP policy {in_dim}-256-256-256-64 ({args.matmul})",
                   "parallelism": f"env-sharded dp{world}", "l2": "rollout + minibatch working set > L2 (inputs "
                   "larger than 126 MB)" if args.workload == "boxworld" else "small working set (latency-bound)"},
        "e2e": {"value": round(e2e_value, 1), "unit": "env-steps/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": round(ms_e2e / args.steps, 3),
                "api": "PPO.train(): host randperm -> pinned H2D per epoch (copy stream, overlapping the previous epoch); "
                       "D2H loss statistics + logger batches into pinned buffers, read by the host after the next "
                       "rollout is enqueued"},
        "gpu_launches": int(launches), "clocks": clk, "replicas_in_sync": in_sync, "roofline": roof,
        "kernel_rooflines": extra,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    if world > 1:
        torch.distributed.destroy_process_group()


class SyntheticProcgen:
    """Host-stepped VecEnv with Procgen's contract (uint8 64x64x3 frames, 15 actions): cycles a pool of pre-generated
    frames, Bernoulli(0.01)*10 rewards, Bernoulli(1/200) dones at no CPU cost."""

    def __init__(self, n, pool=64, seed=0):
        from tpp_b200.discrete_env.pre_vec_env import Box, Discrete
        rng = np.random.default_rng(seed)
        self.n = n
        self.frames = rng.integers(0, 256, (pool, n, 64, 64, 3), dtype=np.uint8)
        self.rew = ((rng.random((pool, n)) < 0.01) * 10.0).astype(np.float32)
        self.done = rng.random((pool, n)) < (1 / 200)
        self.i = 0
        self.observation_space = Box(np.zeros((3, 64, 64)), np.ones((3, 64, 64)))
        self.action_space = Discrete(15)

    def reset(self):
        return self.frames[0]

    def step(self, act):
        self.i = (self.i + 1) % len(self.frames)
        return self.frames[self.i], self.rew[self.i], self.done[self.i], None

    def close(self):
        pass


def run_procgen(args):
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local)
    device = f"cuda:{local}"
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=torch.device(device))
    from tpp_b200 import _lib
    from tpp_b200.agents.ppo import PPO
    from tpp_b200.common.model import ImpalaModel
    from tpp_b200.common.policy import CategoricalPolicy
    from tpp_b200.common.storage import Storage
    hp = WORKLOADS["procgen"]
    N, T = hp["n_envs"], hp["n_steps"]
    matmul = args.matmul if args.matmul in ("tf32x3", "tf32") else "tf32x3"
    env = SyntheticProcgen(N, seed=rank)
    torch.manual_seed(6033)
    pol = CategoricalPolicy(ImpalaModel(3), False, 15).to(device).flatten_()
    st = Storage((3, 64, 64), 256, T, N, device)
    agent = PPO(env, pol, None, st, device, 0, **{k: hp[k] for k in PPO_KEYS}, sample_seed=17 + rank, matmul=matmul)
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
    for _ in range(max(args.warmup, 3)):
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
    st.h2d_bytes = 0
    agent.t = 0
    w0 = time.perf_counter()
    agent.train(T * N * args.steps)
    barrier()
    ms_e2e = (time.perf_counter() - w0) * 1e3
    h2d = st.h2d_bytes // args.steps
    times = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=device)
    if world > 1:
        torch.distributed.all_reduce(times, op=torch.distributed.ReduceOp.MAX)
    ms, ms_e2e = times.tolist()
    if rank != 0:
        if world > 1:
            torch.distributed.destroy_process_group()
        return
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
            "wgrad_us": round(forms["wgrad"], 1),
            "gathered_operand_GBps": round(18 * pixels * rck.SHAPE["C"] * 4 / forms["forward"] / 1e3, 1),
            "note": f"peak = dense bf16 ({pk['source']}); algorithmic FLOPs 2*pixels*9*Cin*Cout; the tile is bound by "
                    "the L2->shared-memory gather of 9 taps x (hi, lo), see profiles/ncu_conv_r01.md"}
    total = N * T * args.steps * world
    mb = min(hp["mini_batch_size"], T * N // hp["n_minibatch"])
    cpu = cpu_baseline("procgen", budget_s=20.0) if not args.no_cpu_baseline else None
    print(json.dumps({
        "metric": "PPO env-steps/sec (rollout+GAE+update)", "value": round(total / (ms * 1e-3), 1),
        "unit": "env-steps/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"procgen-shaped PPO (coinrun hard-500): n_envs={N}/GPU, n_steps={T}, epoch=3, "
                               f"minibatch={mb}, IMPALA-CNN policy ({matmul}), synthetic 64x64x3 uint8 frames",
                   "parallelism": f"env-sharded dp{world}", "l2": "minibatch activations (GBs) >> L2"},
        "e2e": {"value": round(total / (ms_e2e * 1e-3), 1), "unit": "env-steps/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(T * N * 4 + 24 * 20 * 8), "ms_per_step": round(ms_e2e / args.steps, 3),
                "api": "PPO.train() with a host-stepped env: per step pinned uint8 frames H2D, actions D2H"},
        "gpu_launches": int(launches), "clocks": clk, "roofline": roof, "cpu_baseline": cpu}))
    if world > 1:
        torch.distributed.destroy_process_group()


# --------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's numpy/torch path (the reference itself is Python and cannot
# travel to the GPU box; oracle/* restates it and is pinned against it bit-for-bit, see tests/)
# --------------------------------------------------------------------------------------------------

def _cpu_setup(name, n_envs, n_steps):
    from oracle import boxworld as obw
    from oracle import ppo as oppo
    from oracle.prevec import OraclePreVec
    hp = dict(WORKLOADS[name], n_envs=n_envs, n_steps=n_steps)
    torch.manual_seed(6033)
    if name == "procgen":
        rng = np.random.default_rng(0)
        pool = rng.integers(0, 256, (8, n_envs, 64, 64, 3), dtype=np.uint8)
        state = {"i": 0}

        def env_step(a):
            state["i"] = (state["i"] + 1) % len(pool)
            return pool[state["i"]], (rng.random(n_envs) < 0.01) * 10.0, rng.random(n_envs) < (1 / 200)
        tf = lambda f: np.ascontiguousarray(f.transpose(0, 3, 1, 2)).astype(np.float32) / 255.0
        pol = oppo.OraclePolicy(oppo.OracleImpala(3), 15)
        opt = oppo.make_adam(pol, hp["learning_rate"])
        kw = dict(epoch=hp["epoch"], n_minibatch=hp["n_minibatch"], mini_batch_size=hp["mini_batch_size"],
                  grad_clip_norm=hp["grad_clip_norm"], eps_clip=hp["eps_clip"], value_coef=hp["value_coef"],
                  entropy_coef=hp["entropy_coef"])

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
    kw = dict(epoch=hp["epoch"], n_minibatch=hp["n_minibatch"], mini_batch_size=hp["mini_batch_size"],
              grad_clip_norm=hp["grad_clip_norm"], eps_clip=hp["eps_clip"], value_coef=hp["value_coef"],
              entropy_coef=hp["entropy_coef"])

    def iteration(obs):
        obs, _ = oppo.ppo_iteration(env_step, obs, pol, opt, n_steps, n_envs, hp["gamma"], hp["lmbda"],
                                    obs_transform=tf, **kw)
        return obs
    return iteration, obs0


CPU_SAMPLE = {"boxworld": (256, 64), "cartpole": (256, 256), "procgen": (16, 32)}   # bounded (n_envs, n_steps)


def cpu_baseline(name, budget_s=20.0, steps=None):
    """Time the oracle port on the host cores on a bounded sample of the same workload."""
    n_envs, n_steps = CPU_SAMPLE[name]
    iteration, obs = _cpu_setup(name, n_envs, n_steps)
    obs = iteration(obs)                     # warm-up
    t0, k = time.perf_counter(), 0
    while True:
        obs = iteration(obs)
        k += 1
        el = time.perf_counter() - t0
        if (steps is not None and k >= steps) or (steps is None and (el > budget_s or k >= 50)):
            break
    return {"value": round(n_envs * n_steps * k / el, 1), "unit": "env-steps/s", "cores": torch.get_num_threads(),
            "host_cpus": os.cpu_count(), "kind": "port",
            "sample": f"{k} PPO iterations of the oracle port at n_envs={n_envs}, n_steps={n_steps} "
                      f"(same hyperparameters, minibatch clamped); numpy env step is single-threaded, torch uses "
                      f"{torch.get_num_threads()} threads"}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    n_envs, n_steps = CPU_SAMPLE[args.workload]
    iteration, obs = _cpu_setup(args.workload, n_envs, n_steps)
    for _ in range(max(1, min(args.warmup, 2))):
        obs = iteration(obs)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        obs = iteration(obs)
    el = time.perf_counter() - t0
    value = n_envs * n_steps * args.steps / el
    cpu = {"value": round(value, 1), "unit": "env-steps/s", "cores": torch.get_num_threads(), "kind": "port",
           "sample": f"{args.steps} PPO iterations at n_envs={n_envs}, n_steps={n_steps} on the host CPU"}
    print(json.dumps({
        "impl": "reference", "metric": "PPO env-steps/sec (rollout+GAE+update)", "value": round(value, 1),
        "unit": "env-steps/s", "n_gpus": int(os.environ.get("WORLD_SIZE", 1)), "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(el / args.steps * 1e3, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}_env_vec PPO (CPU port of the reference path, bounded sample "
                               f"n_envs={n_envs}, n_steps={n_steps}; hyperparameters of the GPU arm)"},
        "cpu_baseline": cpu,
        "e2e": {"value": round(value, 1), "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="boxworld", choices=list(WORKLOADS))
    ap.add_argument("--matmul", default="tf32x3", choices=["tf32x3", "tf32", "fp32"],
                    help="dense-layer arithmetic: tcgen05 3xTF32 (fp32-parity, default), tcgen05 single TF32, CUDA-core fp32")
    ap.add_argument("--fuse-accum", default="auto",
                    help="minibatches of one gradient-accumulation window sharing a forward/backward pass (auto = all)")
    ap.add_argument("--rollout-chains", type=int, default=1,
                    help="env ranges stepping through the rollout as concurrent kernel chains (experiment: no gain, "
                         "the T sequential steps are latency-bound whatever the range size)")
    ap.add_argument("--no-kernel-rooflines", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU leg (profiling runs only)")
    ap.add_argument("--timed-region-only", action="store_true", help="stop after the device-timed loop (ncu runs)")
    ap.add_argument("--only-kernel-rooflines", action="store_true", help="time just the HBM-bound kernels")
    args = ap.parse_args()
    if args.only_kernel_rooflines:
        print(json.dumps(kernel_rooflines(peaks(), "cuda:0")))
    elif args.impl == "reference":
        run_reference(args)
    elif args.workload == "procgen":
        run_procgen(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
