"""Per-launch timeline of ONE ImpalaEngineTC forward + backward on a Procgen-shaped minibatch (torch.profiler / CUPTI,
warm, not serialised), each launch labelled with the GEMM shape it was called with.

    python profiles/impala_breakdown.py [--mb 2048] [--matmul tf32x3|tf32] > profiles/impala_breakdown_rNN.md
"""
import argparse
import os
import re
import sys
from collections import defaultdict

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

MINE = ("gemm_tc_kernel", "im2col3x3", "maxpool_", "colsum_narrow", "head_backward", "bias_act_split", "split_tf32",
        "conv3x3_wgrad", "conv3x3_fwd_first")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=2048)
    ap.add_argument("--matmul", default="tf32x3")
    ap.add_argument("--per-launch", action="store_true")
    args = ap.parse_args()
    from tpp_b200 import _lib
    from tpp_b200.common.engine import ImpalaEngineTC
    from tpp_b200.common.model import ImpalaModel
    from tpp_b200.common.policy import CategoricalPolicy
    torch.manual_seed(0)
    pol = CategoricalPolicy(ImpalaModel(3), False, 15).to("cuda").flatten_()
    eng = ImpalaEngineTC(pol, 15, (3, 64, 64), precision=3 if args.matmul == "tf32x3" else 1)
    M = args.mb
    x = torch.rand(M, 3 * 64 * 64, device="cuda")
    dhead = torch.randn(M, eng.ld_head, device="cuda") * 1e-3
    labels = []
    real_call = _lib.call

    def logged(name, *a):
        if name == "tpp_gemm_tc":
            g = a[0]._obj
            labels.append(f"gemm M={g.M} N={g.N} K={g.K} mn={g.a_mn}{g.b_mn} split={g.split_k} flags={g.flags}"
                          + (f" conv C={g.conv_C}" if g.conv_C else ""))
        elif name == "tpp_im2col3x3":
            labels.append(f"im2col B={a[2]} {a[3]}x{a[4]}x{a[5]} Kp={a[14]}")
        else:
            labels.append(name)
        return real_call(name, *a)

    def step():
        eng.forward(x, M, train=False)
        eng.backward(dhead, M)

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    _lib.call = logged
    real_try = _lib.try_call

    def logged_try(name, *a):
        ok = real_try(name, *a)
        if ok:
            labels.append(name + (f" cin={a[-3]} cout={a[-2]} W={a[-4]}" if name.startswith("tpp_conv3x3") else ""))
        return ok
    _lib.try_call = logged_try
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        e0.record()
        step()
        e1.record()
        torch.cuda.synchronize()
    _lib.call = real_call
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    mine = [e for e in evs if any(k in e.name for k in MINE)]
    other = sum(e.device_time for e in evs if e not in mine)
    print(f"mb={M} matmul={args.matmul}: forward+backward {e0.elapsed_time(e1):.2f} ms; own kernels "
          f"{sum(e.device_time for e in mine) / 1e3:.2f} ms in {len(mine)} launches, torch glue {other / 1e3:.2f} ms "
          f"(labels {len(labels)})\n")
    tot = defaultdict(lambda: [0, 0.0])
    for e, lab in zip(mine, labels):
        key = re.sub(r"B=\d+ ", "", lab)
        tot[key][0] += 1
        tot[key][1] += e.device_time
    print("| call | launches | total ms | avg us |")
    print("|---|---:|---:|---:|")
    for k, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"| {k} | {n} | {us / 1e3:.3f} | {us / n:.1f} |")
    if args.per_launch:
        print()
        for e, lab in zip(mine, labels):
            print(f"{e.device_time:9.1f} us  {lab}")


if __name__ == "__main__":
    main()
