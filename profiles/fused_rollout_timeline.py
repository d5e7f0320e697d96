"""In-kernel clock64 timeline of the fused rollout-step policy kernel (cluster 0, first epilogue warp of every CTA):
per layer the milestones 0 = enter, 1 = accumulator complete, 2 = peers' inboxes writable, 3 = partials sent,
4 = inbox full, 5 = next operand written.  Also times the kernel alone (CUDA graph of 20 launches)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from tpp_b200.common.engine import MLPEngineTC  # noqa: E402
from tpp_b200.common.model import MLPModel  # noqa: E402
from tpp_b200.common.policy import CategoricalPolicy  # noqa: E402


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    raw = (sys.argv[2] if len(sys.argv) > 2 else "raw") == "raw"
    in_dim, A = (588, 4) if raw else (9, 2)
    torch.manual_seed(0)
    pol = CategoricalPolicy(MLPModel(in_dim, 4, 256, 64), False, A).to("cuda").flatten_()
    eng = MLPEngineTC(pol, A, raw_pixels=raw)
    if raw:
        x = torch.zeros(N, eng.ld_in, device="cuda")
        x[:, :in_dim] = torch.randint(0, 256, (N, in_dim), device="cuda").float()
        ldx = eng.ld_in
    else:
        ldx = (N + 31) // 32 * 32
        x = torch.randn(in_dim, ldx, device="cuda")
    act = torch.zeros(N, dtype=torch.int32, device="cuda")
    logp, value = torch.zeros(N, device="cuda"), torch.zeros(N, device="cuda")
    tick = torch.zeros(1, dtype=torch.int64, device="cuda")
    dbg = torch.zeros(4 * 64, dtype=torch.int64, device="cuda")
    for _ in range(3):
        eng.rollout_fused(x, N, ldx, raw, act, logp, value, 0, tick, 0, dbg=dbg)
    torch.cuda.synchronize()
    d = dbg.view(4, 64).cpu()
    t0 = int(d[:, 32].min())
    print(f"N={N} raw={raw}: cycles since the first CTA's epilogue start (cluster 0)")
    for r in range(4):
        print(f" rank {r}: start {int(d[r, 32]) - t0}")
        for l in range(4):
            print("   layer", l, [int(v) - t0 if int(v) else 0 for v in d[r, l * 8:l * 8 + 8]])
    dt = bench.time_kernel(lambda: eng.rollout_fused(x, N, ldx, raw, act, logp, value, 0, tick, 0), iters=20)
    print(f"kernel alone: {dt * 1e6:.2f} us")


if __name__ == "__main__":
    main()
