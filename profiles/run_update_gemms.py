"""One gather + policy forward + backward pass of the bench's MLP policy at the fused accumulation-window size
(16 x 8192 rows), twice, for `ncu --set full -k regex:gemm_tc_kernel` (the second pass is the warm one):

    python profiles/run_update_gemms.py          # plain run (must exit 0 before the ncu run)
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

hp = bench.workload_hp("boxworld")
_matmul = "tf32x3"
agent, _ = bench.build_agent("boxworld", hp, 0, "cuda:0", matmul=_matmul)
st, env, eng = agent.storage, agent.env, agent.engine
env.reset_rollout(st)
agent.collect_rollout(env, st)
st.compute_estimates(agent.gamma, agent.lmbda, True, True)
M = 16 * 8192
buf = st.minibatch_buffers(M, *agent._obs_buf_args(st))
idx = torch.randperm(st.num_steps * st.num_envs, device="cuda")[:M].contiguous()
dhead = torch.randn(M, eng.ld_head, device="cuda") / 8192
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(2):
    st.gather(idx, buf)
    eng.forward(buf.obs, M, x_lo=buf.obs_lo, raw=buf.raw)
    eng.backward(dhead, M)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok")
