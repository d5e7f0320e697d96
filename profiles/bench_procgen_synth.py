"""BASELINE configs[3]/[4] shape (coinrun / maze / heist hard-500, IMPALA-CNN) with a synthetic host engine.

The Procgen engine is a closed C++ library that is not in this image, so the host side is a stand-in
that hands out pre-generated uint8 64x64x3 frames, Bernoulli(0.01)*10 rewards and Bernoulli(1/200) dones
at no CPU cost: what is measured is this repo's staging + GPU pipeline — pinned H2D of the frames into the uint8
rollout, policy forward + Philox sampling on the device, GAE, gather, fused loss, clip+Adam.  --matmul tf32x3 / tf32
runs the IMPALA convolutions on this repo's im2col + tcgen05 GEMM path (common/engine.py::ImpalaEngineTC), --matmul
library on torch/cuDNN fp32 (TorchModuleEngine) for comparison.

    python profiles/bench_procgen_synth.py [--n-envs 64] [--iters 3] [--matmul tf32x3|tf32|library]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class SyntheticProcgen:
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-envs", type=int, default=64)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--matmul", default="tf32x3", choices=["tf32x3", "tf32", "library"])
    ap.add_argument("--phase-times", action="store_true", help="also time rollout and update separately")
    args = ap.parse_args()
    from tpp_b200.agents.ppo import PPO
    from tpp_b200.common.model import ImpalaModel
    from tpp_b200.common.policy import CategoricalPolicy
    from tpp_b200.common.storage import Storage
    N, T = args.n_envs, 256
    env = SyntheticProcgen(N)
    torch.manual_seed(6033)
    pol = CategoricalPolicy(ImpalaModel(3), False, 15).to("cuda").flatten_()
    st = Storage((3, 64, 64), 256, T, N, "cuda")
    # hard-500 set (hyperparams/procgen/config.yml:81-99): 3 epochs, n_minibatch 8 (default), mini_batch_size 8192
    agent = PPO(env, pol, None, st, "cuda", 0, n_steps=T, n_envs=N, epoch=3, n_minibatch=8, mini_batch_size=8192,
                gamma=0.999, lmbda=0.95, learning_rate=5e-4, entropy_coef=0.01, matmul=args.matmul)
    agent.train(T * N * 2)                   # warm-up: eager iteration, then the iteration that captures the graphs
    torch.cuda.synchronize()
    st.h2d_bytes = 0
    agent.t = 0
    t0 = time.perf_counter()
    agent.train(T * N * args.iters)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out = {"workload": f"procgen-shaped PPO, IMPALA-CNN ({type(agent.engine).__name__}, matmul={args.matmul}), "
                       f"n_envs={N}, n_steps={T}, minibatch={agent.mini_batch_size}, synthetic host frames",
           "env_steps_per_s": round(T * N * args.iters / dt, 1), "s_per_iteration": round(dt / args.iters, 4),
           "h2d_bytes_per_iteration": st.h2d_bytes // args.iters}
    if args.phase_times:
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            agent.optimize()
        torch.cuda.synchronize()
        out["s_per_update_phase"] = round((time.perf_counter() - t0) / 3, 4)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
