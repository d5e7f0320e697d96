"""clock64 timeline of one CTA of gemm_tc_kernel<128> at the update phase's layer shapes: first-wave CTA (y = 0) and a
steady-state CTA (y = 600) of the fused accumulation window (M = 131072)."""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tpp_b200 import _lib
from gemm_overhead import build
names = ["start", "init done", "first TMA issue", "first stage landed", "last MMA committed", "epilogue start",
         "epilogue end", "kernel end", "ep: tmem loaded", "ep: staged", "ep: group 0 written"]
for (M, N, K, prec, outs, y) in [(8192, 256, 256, 3, 1, 0), (8192, 256, 256, 3, 1, 40), (131072, 256, 256, 3, 1, 0),
                                 (131072, 256, 256, 3, 1, 600), (131072, 256, 256, 3, 1, 900),
                                 (131072, 256, 256, 3, -1, 600), (131072, 256, 256, 1, 1, 600)]:
    g, keep = build(M, N, K, prec, 128, outs, 3)
    dbg = torch.zeros(16, dtype=torch.int64, device="cuda")
    g.dbg = dbg.data_ptr()
    g._reserved = y
    for _ in range(3):
        _lib.call("tpp_gemm_tc", C.byref(g), _lib.stream_ptr())
    torch.cuda.synchronize()
    t = dbg.cpu().numpy()
    print(f"M,N,K={M},{N},{K} prec={prec} outs={outs} y={y}: " + ", ".join(f"{n}=+{int(t[i] - t[0])}" for i, n in enumerate(names)))
