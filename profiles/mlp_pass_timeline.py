"""Per-launch timeline of ONE gather + policy forward + loss + backward pass of the bench's MLP policy at M rows
(torch.profiler / CUPTI, warm): which launches the update phase's time goes to, at the minibatch size (8192) and at
the fused accumulation-window size (G x 8192).

    python profiles/mlp_pass_timeline.py [--rows 131072] [--matmul tf32x3]
"""
import argparse
import os
import re
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=131072)
    ap.add_argument("--matmul", default="tf32x3")
    ap.add_argument("--wide-tile-rows", type=int, default=None, help="override MLPEngineTC.wide_tile_rows")
    ap.add_argument("--small-tile-elems", type=int, default=None, help="override MLPEngineTC.small_tile_elems")
    ap.add_argument("--pair-block-n", type=int, default=None, help="512: one tile per CTA pair, 513: persistent pairs")
    args = ap.parse_args()
    hp = bench.workload_hp("boxworld")
    _matmul = args.matmul
    agent, _ = bench.build_agent("boxworld", hp, 0, "cuda:0", matmul=_matmul)
    st, env, eng = agent.storage, agent.env, agent.engine
    if args.wide_tile_rows is not None:
        eng.wide_tile_rows = args.wide_tile_rows
    if os.environ.get("TPP_PAIR_MIN_N"):
        eng.pair_min_n = int(os.environ["TPP_PAIR_MIN_N"])
    if args.pair_block_n is not None:
        eng.pair_block_n = args.pair_block_n
    if args.small_tile_elems is not None:
        eng.small_tile_elems = args.small_tile_elems
    env.reset_rollout(st)
    agent.collect_rollout(env, st)
    st.compute_estimates(agent.gamma, agent.lmbda, True, True)
    M = args.rows
    buf = st.minibatch_buffers(M, *agent._obs_buf_args(st))
    idx = torch.randperm(st.num_steps * st.num_envs, device="cuda")[:M].contiguous()
    dhead = torch.randn(M, eng.ld_head, device="cuda") / 8192

    def one_pass():
        st.gather(idx, buf)
        eng.forward(buf.obs, M, x_lo=buf.obs_lo, raw=buf.raw)
        eng.backward(dhead, M)
    for _ in range(3):
        one_pass()
    torch.cuda.synchronize()
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        one_pass()
        torch.cuda.synchronize()
    evs = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA),
                 key=lambda e: e.time_range.start)
    total = sum(e.device_time for e in evs)
    print(f"rows={M} matmul={args.matmul}: {len(evs)} launches, {total:.1f} us of kernel time "
          f"({total / M * 8192:.1f} us per 8192 rows)\n")
    print("| # | kernel | us |")
    print("|---:|---|---:|")
    for i, e in enumerate(evs):
        name = re.sub(r"^void ", "", re.sub(r"\(.*", "", e.name))
        print(f"| {i} | `{name[:70]}` | {e.device_time:.1f} |")


if __name__ == "__main__":
    main()
