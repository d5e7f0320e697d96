"""Per-kernel time breakdown of ONE PPO iteration of the bench workload with torch.profiler (CUPTI activity tracing:
non-serialised, warm-cache durations — complements the ncu launch list, which is cold-cache and serialised).

    python profiles/kernel_breakdown.py [--workload boxworld|cartpole] [--matmul tf32x3|tf32|fp32] > profiles/breakdown_rNN.md
"""
import argparse
import os
import re
import sys
from collections import defaultdict

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="boxworld")
    ap.add_argument("--matmul", default="tf32x3")
    args = ap.parse_args()
    hp = bench.workload_hp(args.workload)
    _matmul = args.matmul
    agent, _ = bench.build_agent(args.workload, hp, 0, "cuda:0", matmul=_matmul)
    st, env = agent.storage, agent.env
    env.reset_rollout(st)

    def iteration():
        agent.collect_rollout(env, st)
        st.compute_estimates(agent.gamma, agent.lmbda, True, True)
        agent.optimize()
        agent._carry_over(st)
    for _ in range(3):
        iteration()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        e0.record()
        iteration()
        e1.record()
        torch.cuda.synchronize()
    wall = e0.elapsed_time(e1)
    tot = defaultdict(lambda: [0, 0.0])
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            name = re.sub(r"\(.*", "", ev.name)
            name = re.sub(r"^void ", "", name)
            tot[name][0] += 1
            tot[name][1] += ev.device_time
    total = sum(v[1] for v in tot.values())
    print(f"workload={args.workload} matmul={args.matmul}: iteration {wall:.2f} ms (CUDA events), "
          f"sum of kernel time {total / 1e3:.2f} ms, {sum(v[0] for v in tot.values())} launches\n")
    print("| kernel | launches | total ms | share of kernel time | avg us |")
    print("|---|---:|---:|---:|---:|")
    for name, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{name[:90]}` | {n} | {us / 1e3:.3f} | {100 * us / total:.1f}% | {us / n:.2f} |")


if __name__ == "__main__":
    main()
