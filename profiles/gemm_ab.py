"""Graph-timed tpp_gemm_tc on the bench MLP's layer shapes (minibatch 8192, 588-256-256-256-64, 3xTF32).
Set TPP_B200_LIB to time another build of the library (A/B comparisons of kernel changes)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tpp_b200 import _lib  # noqa: E402


def build(M, N, K, a_mn=0, b_mn=0, split_k=1, flags=3, block_n=0, pair=True, plain=False, mask=False, colsum=False):
    lda = (M + 31) // 32 * 32 if a_mn else (K + 31) // 32 * 32
    ldb = (N + 31) // 32 * 32 if b_mn else (K + 31) // 32 * 32
    ra, rb = ((K, lda) if a_mn else (M, lda)), ((K, ldb) if b_mn else (N, ldb))
    keep = dict(a=[torch.randn(*ra, device="cuda") for _ in range(2)], b=[torch.randn(*rb, device="cuda") for _ in range(2)])
    ldc = (N + 31) // 32 * 32
    keep["out"] = [torch.zeros(M, ldc, device="cuda") for _ in range(3)]
    keep["bias"], keep["mask"], keep["cs"] = torch.zeros(N, device="cuda"), torch.randn(M, ldc, device="cuda"), torch.zeros(N, device="cuda")
    g = _lib.TcGemm()
    g.a_hi, g.a_lo, g.lda = keep["a"][0].data_ptr(), keep["a"][1].data_ptr(), lda
    g.b_hi, g.b_lo, g.ldb = keep["b"][0].data_ptr(), keep["b"][1].data_ptr(), ldb
    g.M, g.N, g.K, g.a_mn, g.b_mn = M, N, K, a_mn, b_mn
    g.precision, g.split_k, g.flags, g.block_n = 3, split_k, flags | (4 if mask else 0), block_n
    g.bias = keep["bias"].data_ptr()
    if mask:
        g.mask, g.ld_mask = keep["mask"].data_ptr(), ldc
    if colsum:
        g.colsum = keep["cs"].data_ptr()
    if flags & 8 or plain:
        g.out = keep["out"][0].data_ptr()
    if pair and not flags & 8:
        g.out_hi, g.out_lo = keep["out"][1].data_ptr(), keep["out"][2].data_ptr()
    g.ldc = ldc
    return g, keep


def timed(g, n=50):
    s = _lib.stream_ptr()
    for _ in range(3):
        _lib.call("tpp_gemm_tc", C.byref(g), s)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for _ in range(n):
            _lib.call("tpp_gemm_tc", C.byref(g), _lib.stream_ptr())
    graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    graph.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


CASES = [
    ("fwd L1 8192x256x588", dict(M=8192, N=256, K=588)),
    ("fwd L2 8192x256x256", dict(M=8192, N=256, K=256)),
    ("fwd L4 8192x64x256 +plain", dict(M=8192, N=64, K=256, flags=1, plain=True)),
    ("head 8192x16x64", dict(M=8192, N=16, K=64, flags=1, pair=False, plain=True, block_n=16)),
    ("dgrad L4 8192x256x64", dict(M=8192, N=256, K=64, b_mn=1, flags=0, mask=True, colsum=True)),
    ("dgrad L3 8192x256x256", dict(M=8192, N=256, K=256, b_mn=1, flags=0, mask=True, colsum=True)),
    ("dgrad L3, no colsum", dict(M=8192, N=256, K=256, b_mn=1, flags=0, mask=True, colsum=False)),
    ("dgrad L3, no mask", dict(M=8192, N=256, K=256, b_mn=1, flags=0, mask=False, colsum=True)),
    ("dgrad L3, neither", dict(M=8192, N=256, K=256, b_mn=1, flags=0)),
    ("dgrad L3, neither, K-major B", dict(M=8192, N=256, K=256, b_mn=0, flags=0)),
    ("wgrad L1 256x588x8192", dict(M=256, N=588, K=8192, a_mn=1, b_mn=1, flags=8, split_k=15, block_n=128)),
    ("wgrad L2 256x256x8192", dict(M=256, N=256, K=8192, a_mn=1, b_mn=1, flags=8, split_k=37, block_n=128)),
    ("wgrad L4 64x256x8192", dict(M=64, N=256, K=8192, a_mn=1, b_mn=1, flags=8, split_k=74, block_n=128)),
]

if __name__ == "__main__":
    print("library:", _lib.LIB_PATH)
    tot = 0.0
    for name, kw in CASES:
        g, keep = build(**kw)
        us = timed(g)
        tot += us
        print(f"{name:32s} {us:8.2f} us")
    print(f"{'sum':32s} {tot:8.2f} us")
