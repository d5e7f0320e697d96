import sys, torch
sys.path.insert(0, '/root/repo')
import bench
from tpp_b200.discrete_env.acrobot_pre_vec import AcrobotVecEnv
from tpp_b200.discrete_env.lunar_lander_pre_vec import LunarLanderVecEnv
N = 1 << 22
for cls, B in ((AcrobotVecEnv, 137), (LunarLanderVecEnv, 32 + 32 + 4 + 8 + 4 + 1)):
    for hint in (4, 2, 1):
        env = cls(n_envs=N, seed=1)
        env._cfg.p[7] = float(hint)
        act = torch.randint(0, env.n_actions, (N,), device="cuda", dtype=torch.int32)
        st = {"c": 0}
        def step():
            c = st["c"]; env.step_into(env._slots[c], env._slots[c ^ 1], act, env._rew, env._done); st["c"] = c ^ 1
        dt = bench.time_kernel(step, iters=30)
        print(cls.__name__, "vec", hint, f"{dt*1e6:.1f} us  {B*N/dt/1e9:.0f} GB/s  {N/dt/1e9:.2f} G env-steps/s")
        del env
