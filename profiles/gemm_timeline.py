"""clock64 timeline of CTA (0,0,0) of gemm_tc_kernel (probe points in csrc/gemm_tc.cu)."""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tpp_b200 import _lib
from gemm_overhead import build
names = ["start", "init done", "first TMA issue", "first stage landed", "last MMA committed", "epilogue start",
         "epilogue end", "kernel end", "ep: tmem loaded", "ep: staged", "ep: group 0 written"]
for (M, N, K, prec, outs) in [(128, 128, 32, 3, 0), (128, 128, 32, 3, 1), (128, 128, 588, 3, 1), (8192, 256, 588, 3, 1),
                              (8192, 256, 256, 1, 1)]:
    g, keep = build(M, N, K, prec, 128, outs, 3)
    dbg = torch.zeros(16, dtype=torch.int64, device="cuda")
    g.dbg = dbg.data_ptr()
    for _ in range(3):
        _lib.call("tpp_gemm_tc", C.byref(g), _lib.stream_ptr())
    torch.cuda.synchronize()
    t = dbg.cpu().numpy()
    print(f"M,N,K={M},{N},{K} prec={prec} outs={outs}: " + ", ".join(f"{n}=+{int(t[i] - t[0])}" for i, n in enumerate(names)))
