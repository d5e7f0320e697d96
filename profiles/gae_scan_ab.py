"""GAE at the PPO configs' sizes: the exact kernels (tpp_gae + tpp_adv_normalize) against the fused warp-level segmented
scan (tpp_gae_scan), graph-timed like bench.py's kernel_rooflines (the moments memset is part of every variant).

    python profiles/gae_scan_ab.py            # TPP_GAE_COOP=0: plain launch instead of the cooperative one (the default)
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from tpp_b200.common.storage import Storage  # noqa: E402


def main():
    print(f"TPP_GAE_COOP={os.environ.get('TPP_GAE_COOP', '1')}")
    print("| T | n_envs | exact gae + normalize | exact gae only | scan, raw adv (no barrier) | scan fused (barrier + normalize) |")
    print("|---:|---:|---:|---:|---:|---:|")
    for T, N in ((256, 256), (256, 1024), (256, 4096), (128, 4096), (64, 4096)):
        st = Storage((1,), 1, T, N, "cuda")
        st.rew.normal_(); st.value.normal_()
        st.done_u8.copy_((torch.rand(T, st.ld, device="cuda") < 0.02).to(torch.uint8))
        row = []
        for mode, norm in (("exact", True), ("exact", False), ("warp_scan", False), ("warp_scan", True)):
            st.gae_mode = mode
            dt = bench.time_kernel(lambda: st.compute_estimates(0.99, 0.95, True, norm), iters=20)
            row.append(f"{dt * 1e6:.2f} us")
        print(f"| {T} | {N} | " + " | ".join(row) + " |")


if __name__ == "__main__":
    main()
