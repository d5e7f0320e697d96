"""Micro-benchmark of tpp_gemm_tc over tile / precision variants (CUDA events, 50 launches each)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tpp_b200 import _lib  # noqa: E402


def run(M, N, K, precision, block_n, a_mn=0, b_mn=0, split_k=1, outs=2, flags=3):
    lda = (M + 31) // 32 * 32 if a_mn else (K + 3) // 4 * 4
    ldb = (N + 31) // 32 * 32 if b_mn else (K + 3) // 4 * 4
    ra, rb = (K, lda) if a_mn else (M, lda), (K, ldb) if b_mn else (N, ldb)
    a = [torch.randn(*ra, device="cuda") for _ in range(2)]
    b = [torch.randn(*rb, device="cuda") for _ in range(2)]
    ldc = (N + 31) // 32 * 32
    out = [torch.zeros(M, ldc, device="cuda") for _ in range(3)]
    bias = torch.zeros(N, device="cuda")
    g = _lib.TcGemm()
    g.a_hi, g.a_lo, g.lda = a[0].data_ptr(), a[1].data_ptr(), lda
    g.b_hi, g.b_lo, g.ldb = b[0].data_ptr(), b[1].data_ptr(), ldb
    g.M, g.N, g.K, g.a_mn, g.b_mn = M, N, K, a_mn, b_mn
    g.precision, g.split_k, g.flags, g.block_n = precision, split_k, flags, block_n
    g.bias = bias.data_ptr()
    if flags & 8:
        g.out = out[0].data_ptr()
    elif outs == 2:
        g.out_hi, g.out_lo = out[1].data_ptr(), out[2].data_ptr()
    elif outs == 1:
        g.out = out[0].data_ptr()
    g.ldc = ldc
    s = _lib.stream_ptr()
    for _ in range(5):
        _lib.call("tpp_gemm_tc", C.byref(g), s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        _lib.call("tpp_gemm_tc", C.byref(g), s)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 50 * 1e3
    return us, 2.0 * M * N * K / us / 1e6


if __name__ == "__main__":
    print("| shape (M,N,K) | kind | precision | block_n | split_k | us | TFLOP/s (2MNK) |")
    print("|---|---|---:|---:|---:|---:|---:|")
    for (M, N, K) in [(8192, 256, 588), (8192, 256, 256), (8192, 64, 256), (4096, 256, 588)]:
        for prec in (3, 1):
            for bn in (64, 128, 256):
                if bn > max(N, 64):
                    continue
                us, tf = run(M, N, K, prec, bn)
                print(f"| {M},{N},{K} | fwd K,K | {prec} | {bn} | 1 | {us:.1f} | {tf:.1f} |")
    for (M, N, K) in [(8192, 256, 256), (8192, 588, 256)]:
        for prec in (3, 1):
            us, tf = run(M, N, K, prec, 128, b_mn=1, flags=4 * 0)
            print(f"| {M},{N},{K} | dgrad K,MN | {prec} | 128 | 1 | {us:.1f} | {tf:.1f} |")
    for (M, N, K, sk) in [(256, 256, 8192, 37), (256, 588, 8192, 15), (256, 256, 8192, 8), (256, 588, 8192, 32),
                          (64, 256, 8192, 74)]:
        for prec in (3, 1):
            us, tf = run(M, N, K, prec, 128, a_mn=1, b_mn=1, split_k=sk, flags=8)
            print(f"| {M},{N},{K} | wgrad MN,MN | {prec} | 128 | {sk} | {us:.1f} | {tf:.1f} |")
