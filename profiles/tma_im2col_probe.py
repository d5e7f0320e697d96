"""Exploration of the TMA im2col conventions on this driver (prints what each box actually contains)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tpp_b200 import _lib as L  # noqa: E402


def make(B, H, W, C):
    b, h, w, c = torch.meshgrid(torch.arange(B), torch.arange(H), torch.arange(W), torch.arange(C), indexing="ij")
    return ((b + 1) * 1000000 + h * 10000 + w * 100 + c).float().cuda().contiguous()


def expected(x, cpp, pixels, w, h, n, ow, oh):
    B, H, W, C = x.shape
    out = torch.zeros(pixels, cpp)
    xc = x.cpu()
    q, p, b = w + 1, h + 1, n          # output pixel of the base
    for i in range(pixels):
        if b < B:
            sy, sx = p - 1 + oh, q - 1 + ow
            if 0 <= sy < H and 0 <= sx < W:
                out[i, :min(C, cpp)] = xc[b, sy, sx, :min(C, cpp)]
        q += 1
        if q == W:
            q, p = 0, p + 1
            if p == H:
                p, b = 0, b + 1
    return out


def run(name, B, H, W, C, cpp, pixels, w, h, n, ow, oh, swz=0):
    x = make(B, H, W, C)
    out = torch.full((pixels, cpp), -1.0, device="cuda")
    try:
        import ctypes
        probe = ctypes.CDLL(os.path.join(ROOT, "tests", "native", "libtpp_probe.so"))     # make -C tests/native
        probe.tpp_debug_tma_im2col.argtypes = [ctypes.c_void_p] + [ctypes.c_int32] * 12 + [ctypes.c_void_p] * 2
        assert probe.tpp_debug_tma_im2col(L.ptr(x), B, H, W, C, cpp, pixels, w, h, n, ow, oh, swz, L.ptr(out),
                                          L.stream_ptr()) == 0
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        print(f"{name}: ERROR {e}")
        return
    got = out.cpu()
    exp = expected(x, cpp, pixels, w, h, n, ow, oh)
    if swz == 128:   # undo the 128B swizzle: 16-byte chunk index ^= (row % 8), rows of 128 B
        g = got.view(pixels, cpp // 4, 4)
        un = torch.zeros_like(g)
        for r in range(pixels):
            for ch in range(cpp // 4):
                un[r, ch] = g[r, ch ^ (r % 8)]
        got = un.view(pixels, cpp)
    ok = torch.equal(got, exp)
    print(f"{name}: match={ok}")
    if not ok:
        bad = (got != exp).any(1).nonzero().flatten()
        print("   first mismatching pixels:", bad[:10].tolist(), "of", len(bad))
        for i in bad[:6].tolist():
            print(f"   pixel {i}: got ch0..3 {got[i, :4].tolist()} ch16..17 {got[i, 16:18].tolist() if cpp > 17 else ''}"
                  f" expected {exp[i, :4].tolist()}")


run("1 C32 base(-1,-1,0) tap(0,0)", 3, 8, 8, 32, 32, 128, -1, -1, 0, 0, 0)
run("2 C32 tap(1,1)", 3, 8, 8, 32, 32, 128, -1, -1, 0, 1, 1)
run("3 C32 tap(2,2)", 3, 8, 8, 32, 32, 128, -1, -1, 0, 2, 2)
run("3b C32 tap(2,0) (w=2,h=0)", 3, 8, 8, 32, 32, 128, -1, -1, 0, 2, 0)
run("4 C32 start (n=1,h=5,w=3) tap(1,1)", 3, 8, 8, 32, 32, 128, 2, 4, 1, 1, 1)
run("5 C32 tail beyond last image", 3, 8, 8, 32, 32, 128, -1, 5, 2, 1, 1)
run("6 C16 cpp32 (channel OOB fill)", 3, 8, 8, 16, 32, 128, -1, -1, 0, 1, 1)
run("7 C16 cpp16", 3, 8, 8, 16, 16, 128, -1, -1, 0, 1, 1)
run("8 C32 swizzle128", 3, 8, 8, 32, 32, 128, 2, 4, 1, 0, 2, 128)
run("9 C4 cpp32", 3, 8, 8, 4, 32, 128, -1, -1, 0, 1, 1)
run("10 C16 cpp32 swizzle128 odd size 7x5", 4, 7, 5, 16, 32, 128, 1, 2, 0, 2, 1, 128)
run("11 C32 32 pixels ATOM_32B", 3, 8, 8, 32, 32, 32, -1, -1, 0, 1, 1, 1)
