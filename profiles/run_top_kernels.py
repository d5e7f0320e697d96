"""Launch each top kernel of the hot path a few times at its bench shape (for `ncu --set full -k regex:...`).

    python profiles/run_top_kernels.py           # plain run (must exit 0 before the ncu run)
"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "profiles"))
from gemm_overhead import build  # noqa: E402
from tpp_b200 import _lib  # noqa: E402
from tpp_b200.common.storage import Storage  # noqa: E402
from tpp_b200.discrete_env.acrobot_pre_vec import AcrobotVecEnv  # noqa: E402
from tpp_b200.discrete_env.cartpole_pre_vec import CartPoleVecEnv  # noqa: E402

N = 1 << 22
for cls in (CartPoleVecEnv, AcrobotVecEnv):
    env = cls(n_envs=N, seed=1)
    act = torch.randint(0, env.n_actions, (N,), device="cuda", dtype=torch.int32)
    for i in range(4):
        env.step_into(env._slots[i & 1], env._slots[(i & 1) ^ 1], act, env._rew, env._done)
    torch.cuda.synchronize()
    del env
T, Ng = 256, 1 << 16
st = Storage((1,), 1, T, Ng, "cuda")
st.rew.normal_(); st.value.normal_()
for _ in range(3):
    st.compute_estimates(0.99, 0.95, True, True)
torch.cuda.synchronize()
g, keep = build(8192, 256, 588, 3, 128, 1, 3)        # layer-1 forward of the bench workload, 3xTF32
for _ in range(4):
    _lib.call("tpp_gemm_tc", C.byref(g), _lib.stream_ptr())
torch.cuda.synchronize()
print("ok")
