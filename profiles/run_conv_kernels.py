"""Launch the three implicit-GEMM convolution forms at the IMPALA block-1 shape of the Procgen workload
(2048 samples x 32x32 pixels, 16 -> 16 channels, 3xTF32), for `ncu --set full -k regex:gemm_tc_kernel`.

    python profiles/run_conv_kernels.py          # plain run first (must exit 0), prints graph-timed durations
    ncu --set full --clock-control none -k regex:gemm_tc_kernel -s 2 -c 1 -o X python profiles/run_conv_kernels.py forward
"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tpp_b200 import _lib as L  # noqa: E402


def pair(x):
    hi = ((x.contiguous().view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)
    return hi, x - hi


SHAPE = dict(B=2048, H=32, W=32, C=16)


def time_forms(only=None):
    """Graph-timed duration (us) of each convolution form; with `only`, three plain launches of that form (ncu)."""
    B, H, W, Cc = SHAPE["B"], SHAPE["H"], SHAPE["W"], SHAPE["C"]
    rows = B * H * W
    torch.manual_seed(0)
    x = pair(torch.randn(B, H, W, Cc, device="cuda").clamp_min(0))
    dy = pair(torch.randn(rows, Cc, device="cuda"))
    w16 = pair(torch.randn(Cc, 144, device="cuda") * 0.1)
    bias = torch.zeros(Cc, device="cuda")
    out = torch.zeros(rows, Cc, device="cuda")
    oh, ol = torch.zeros_like(out), torch.zeros_like(out)
    mask = torch.randn(rows, Cc, device="cuda")
    gw = torch.zeros(288, Cc, device="cuda")
    cs = torch.zeros(Cc, device="cuda")

    fwd = L.TcGemm()
    fwd.a_hi, fwd.a_lo, fwd.b_hi, fwd.b_lo, fwd.ldb = x[0].data_ptr(), x[1].data_ptr(), w16[0].data_ptr(), w16[1].data_ptr(), 144
    fwd.M, fwd.N, fwd.K, fwd.precision, fwd.split_k = rows, Cc, 144, 3, 1
    fwd.conv_B, fwd.conv_H, fwd.conv_W, fwd.conv_C = B, H, W, Cc
    fwd.flags, fwd.bias = L.EPI_BIAS | L.EPI_PAIR_RELU, bias.data_ptr()
    fwd.out, fwd.out_hi, fwd.out_lo, fwd.ldc = out.data_ptr(), oh.data_ptr(), ol.data_ptr(), Cc

    dg = L.TcGemm()
    dg.a_hi, dg.a_lo, dg.b_hi, dg.b_lo, dg.ldb = dy[0].data_ptr(), dy[1].data_ptr(), w16[0].data_ptr(), w16[1].data_ptr(), 144
    dg.M, dg.N, dg.K, dg.precision, dg.split_k = rows, Cc, 144, 3, 1
    dg.conv_B, dg.conv_H, dg.conv_W, dg.conv_C = B, H, W, Cc
    dg.flags, dg.mask, dg.ld_mask, dg.addend, dg.ld_add = L.EPI_MASK | L.EPI_ADD, mask.data_ptr(), Cc, mask.data_ptr(), Cc
    dg.out, dg.out_hi, dg.out_lo, dg.ldc, dg.colsum = out.data_ptr(), oh.data_ptr(), ol.data_ptr(), Cc, cs.data_ptr()

    wg = L.TcGemm()
    wg.a_hi, wg.a_lo, wg.b_hi, wg.b_lo, wg.ldb = x[0].data_ptr(), x[1].data_ptr(), dy[0].data_ptr(), dy[1].data_ptr(), Cc
    wg.M, wg.N, wg.K, wg.precision, wg.split_k, wg.a_mn, wg.b_mn = 288, Cc, rows, 3, rows // 1024, 1, 1
    wg.conv_B, wg.conv_H, wg.conv_W, wg.conv_C, wg.conv_wgrad = B, H, W, Cc, 1
    wg.flags, wg.out, wg.ldc, wg.block_n = L.EPI_ACCUM, gw.data_ptr(), Cc, 32

    x_plain = (x[0] + x[1]).contiguous()
    dy_plain = (dy[0] + dy[1]).contiguous()

    wf_plain = (w16[0] + w16[1]).contiguous()

    def launch(name, g):
        if name == "forward_fma":   # csrc/conv_cc.cu: same epilogue as `forward` (bias, plain + TF32 pair of relu(y))
            L.call("tpp_conv3x3_fma", L.ptr(x_plain), 1, L.ptr(wf_plain), 16, L.ptr(bias), None, None, 1, L.ptr(out),
                   L.ptr(oh), L.ptr(ol), None, B, H, W, Cc, Cc, L.stream_ptr())
        elif name == "dgrad_fma":   # ... and as `dgrad` (ReLU mask, skip-gradient add, bias-gradient column sums)
            L.call("tpp_conv3x3_fma", L.ptr(dy_plain), 0, L.ptr(wf_plain), 16, None, L.ptr(mask), L.ptr(mask), 0,
                   L.ptr(out), L.ptr(oh), L.ptr(ol), L.ptr(cs), B, H, W, Cc, Cc, L.stream_ptr())
        elif name == "wgrad_fma":     # csrc/conv_cc.cu: the weight gradient on the fp32 FMA pipe (the engine's default)
            L.call("tpp_conv3x3_wgrad", L.ptr(x_plain), 1, L.ptr(dy_plain), L.ptr(gw), B, H, W, Cc, Cc, L.stream_ptr())
        else:
            L.call("tpp_gemm_tc", C.byref(g), L.stream_ptr())

    res = {}
    for name, g in (("forward", fwd), ("dgrad", dg), ("wgrad", wg), ("forward_fma", None), ("dgrad_fma", None),
                    ("wgrad_fma", None)):
        if only:
            if name == only:
                for _ in range(3):
                    launch(name, g)
                torch.cuda.synchronize()
            continue
        for _ in range(2):
            launch(name, g)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for _ in range(10):
                launch(name, g)
        graph.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        graph.replay()
        e1.record()
        torch.cuda.synchronize()
        res[name] = e0.elapsed_time(e1) / 10 * 1e3
    return res


def main():
    only = sys.argv[1] if len(sys.argv) > 1 else None      # forward | dgrad | wgrad | wgrad_fma: 3 plain launches (ncu: -s 2 -c 1)
    res = time_forms(only)
    B, H, W, Cc = SHAPE["B"], SHAPE["H"], SHAPE["W"], SHAPE["C"]
    rows = B * H * W
    flops = 2.0 * rows * Cc * 9 * Cc
    act_bytes = rows * Cc * 4
    for name, us in res.items():
        print(f"{name}: {us:.1f} us  = {flops / us / 1e6:.1f} TFLOP/s algorithmic (2*pixels*9*Cin*Cout), "
              f"gathered operand bytes 9 taps x 2 (hi, lo) x {act_bytes / 1e6:.0f} MB = {18 * act_bytes / us / 1e3:.0f} GB/s")


if __name__ == "__main__":
    main()
