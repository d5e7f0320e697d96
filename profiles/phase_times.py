"""CUDA-event time of the three phases of one PPO iteration of the bench workload (rollout / GAE / update), averaged
over a few warm iterations.

    python profiles/phase_times.py [--workload boxworld] [--iters 5]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="boxworld")
    ap.add_argument("--matmul", default="tf32x3")
    ap.add_argument("--iters", type=int, default=5)
    args = ap.parse_args()
    hp = bench.workload_hp(args.workload)
    _matmul = args.matmul
    agent, _ = bench.build_agent(args.workload, hp, 0, "cuda:0", matmul=_matmul)
    st, env = agent.storage, agent.env
    env.reset_rollout(st)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.iters + 3)]
    for it in range(args.iters + 3):
        ev[it][0].record()
        agent.collect_rollout(env, st)
        ev[it][1].record()
        st.compute_estimates(agent.gamma, agent.lmbda, True, True)
        ev[it][2].record()
        agent.optimize()
        ev[it][3].record()
        agent._carry_over(st)
    torch.cuda.synchronize()
    ph = [sum(e[i].elapsed_time(e[i + 1]) for e in ev[3:]) / args.iters for i in range(3)]
    T = st.num_steps
    print(f"workload={args.workload}: rollout {ph[0]:.2f} ms ({ph[0] / T * 1e3:.1f} us/step), GAE {ph[1]:.3f} ms, "
          f"update {ph[2]:.2f} ms, sum {sum(ph):.2f} ms")


if __name__ == "__main__":
    main()
