"""One eager (graph-free) PPO iteration of the bench workload for `ncu -k regex:... -c N` captures (round 2):
2 rollouts (fused_policy_kernel, boxworld_step / reset kernels), GAE, one update epoch (gather_img, gemm_tc pair kernels,
head_backward, ppo_loss, grad_sqnorm / adam_clip).

    python profiles/run_round2_kernels.py            # must exit 0 before any ncu run
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    hp = bench.workload_hp("boxworld")
    agent, _ = bench.build_agent("boxworld", hp, 0, "cuda:0", use_cuda_graph=False)
    agent.epoch = 1
    st, env = agent.storage, agent.env
    env.reset_rollout(st)
    for _ in range(2):
        agent.collect_rollout(env, st)
        st.compute_estimates(agent.gamma, agent.lmbda, True, True)
        agent._carry_over(st)
    agent.optimize()          # first window eager-allocates; profile filters skip into the second call with -s
    agent.optimize()
    torch.cuda.synchronize()
    print("ok")


if __name__ == "__main__":
    main()
