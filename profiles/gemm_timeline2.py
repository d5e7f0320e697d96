import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tpp_b200 import _lib
from gemm_overhead import build
for (flags, outs) in [(0, 0), (1, 0), (2, 0), (3, 0), (0, 1), (3, 1)]:
    g, keep = build(128, 128, 32, 3, 128, outs, flags)
    dbg = torch.zeros(16, dtype=torch.int64, device="cuda")
    g.dbg = dbg.data_ptr()
    for _ in range(3):
        _lib.call("tpp_gemm_tc", C.byref(g), _lib.stream_ptr())
    torch.cuda.synchronize()
    t = dbg.cpu().numpy()
    print(f"flags={flags} outs={outs}: staged->group0 done = {int(t[10]-t[9])} cycles, epilogue total = {int(t[6]-t[5])}; "
          f"it0 +{int(t[11]-t[9])} it1 +{int(t[12]-t[9])} it2 +{int(t[13]-t[9])} it3 +{int(t[14]-t[9])}")
