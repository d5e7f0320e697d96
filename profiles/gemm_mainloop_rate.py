"""Steady-state rate of the 256 x 256 CTA-pair GEMM main loop: one tile per pair, K = 4096 (128 k-blocks), so the time is
all TMA -> (split) -> MMA.  Pair operands from HBM vs lo halves formed on chip, and the layer shape (K = 256)."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tpp_b200 import _lib  # noqa: E402


def run(M, N, K, prec, block_n, out_mode, reps=6):
    a = torch.randn(M, K, device="cuda")
    b = torch.randn(N, K, device="cuda") * 0.05
    a2, b2 = a.clone(), b.clone()
    out = torch.zeros(M, N, device="cuda")
    out2 = torch.zeros(M, N, device="cuda")
    g = _lib.TcGemm()
    g.a_hi, g.a_lo, g.lda = a.data_ptr(), a2.data_ptr(), K
    g.b_hi, g.b_lo, g.ldb = b.data_ptr(), b2.data_ptr(), K
    g.M, g.N, g.K, g.precision, g.block_n, g.ldc = M, N, K, prec, block_n, N
    if out_mode == "plain":
        g.out = out.data_ptr()
    else:
        g.out_hi, g.out_lo = out.data_ptr(), out2.data_ptr()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.call("tpp_gemm_tc", C.byref(g), _lib.stream_ptr())
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return sorted(ts)[len(ts) // 2]


def main():
    S = _lib.TC_A_SPLIT | _lib.TC_B_SPLIT
    for name, M, N, K in (("one tile per pair, K=4096", 256 * 74, 256, 4096), ("layer shape, K=256", 131072, 256, 256),
                          ("two tiles per pair, K=2048", 256 * 148, 256, 2048)):
        for tag, prec, om in (("pairs -> pair out", 3, "pair"), ("pairs -> plain out", 3, "plain"),
                              ("A split on chip -> plain", 3 | _lib.TC_A_SPLIT, "plain"),
                              ("A+B split on chip -> plain", 3 | S, "plain"), ("1xTF32 -> plain", 1, "plain")):
            try:
                t = run(M, N, K, prec, int(os.environ.get("TILE", "513")), om)
            except Exception as e:
                print(f"{name:28s} {tag:28s} {str(e)[-60:]}")
                continue
            kb = (M // 256) * (K // 32) / 74.0
            print(f"{name:28s} {tag:28s} {t:8.1f} us   {t / kb * 1e3:7.1f} ns / k-block / pair   "
                  f"{2.0 * M * N * K / t * 1e-6:7.1f} TFLOP/s")


if __name__ == "__main__":
    main()
