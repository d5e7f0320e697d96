"""Where does the fixed cost of gemm_tc go?  GPU-side time per launch from a CUDA graph of 20 back-to-back launches
(no CPU launch overhead in the number)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tpp_b200 import _lib  # noqa: E402


def build(M, N, K, precision, block_n, outs, flags):
    lda = ldb = (K + 3) // 4 * 4
    a = [torch.randn(M, lda, device="cuda") for _ in range(2)]
    b = [torch.randn(N, ldb, device="cuda") for _ in range(2)]
    ldc = (N + 31) // 32 * 32
    out = [torch.zeros(M, ldc, device="cuda") for _ in range(3)]
    bias = torch.zeros(N, device="cuda")
    g = _lib.TcGemm()
    g.a_hi, g.a_lo, g.lda = a[0].data_ptr(), a[1].data_ptr(), lda
    g.b_hi, g.b_lo, g.ldb = b[0].data_ptr(), b[1].data_ptr(), ldb
    g.M, g.N, g.K = M, N, K
    g.precision, g.split_k, g.flags, g.block_n = precision, 1, flags, block_n
    g.bias = bias.data_ptr()
    if outs >= 1:
        g.out_hi, g.out_lo = out[1].data_ptr(), out[2].data_ptr()
    if outs == 3 or outs == -1:
        g.out = out[0].data_ptr()
    if outs == -1:
        g.out_hi = g.out_lo = None
    g.ldc = ldc
    return g, (a, b, out, bias)


def time_graph(g, reps=20):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            _lib.call("tpp_gemm_tc", C.byref(g), _lib.stream_ptr())
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for _ in range(reps):
                _lib.call("tpp_gemm_tc", C.byref(g), _lib.stream_ptr())
        graph.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            graph.replay()
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (5 * reps) * 1e3


if __name__ == "__main__":
    print("| M,N,K | prec | block_n | outputs | us/launch (graph) |")
    print("|---|---:|---:|---|---:|")
    for (M, N, K) in [(8192, 256, 32), (8192, 256, 256), (8192, 256, 588), (128, 128, 32), (128, 128, 588)]:
        for prec in (3, 1):
            for outs, name in ((0, "none"), (-1, "plain"), (1, "hi+lo"), (3, "plain+hi+lo")):
                g, keep = build(M, N, K, prec, 128, outs, 3)
                print(f"| {M},{N},{K} | {prec} | 128 | {name} | {time_graph(g):.2f} |", flush=True)
