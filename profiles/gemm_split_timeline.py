"""clock64 timeline of the on-chip operand splitters of one CTA pair (gemm_tc_kernel, TPP_TC_A_SPLIT | TPP_TC_B_SPLIT):
per k-block, when the stage landed, when the lo halves were written, fenced, signalled, and when the MMA issuer saw it."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tpp_b200 import _lib  # noqa: E402

M, N, K = 256 * 74, 256, 4096
a = torch.randn(M, K, device="cuda")
b = torch.randn(N, K, device="cuda") * 0.05
out = torch.zeros(M, N, device="cuda")
for bn in (513, 512, 256, 128):
    g = _lib.TcGemm()
    g.a_hi, g.lda, g.b_hi, g.ldb = a.data_ptr(), K, b.data_ptr(), K
    g.M, g.N, g.K, g.block_n, g.ldc, g.out = M, N, K, bn, N, out.data_ptr()
    g.precision = 3 | _lib.TC_A_SPLIT | _lib.TC_B_SPLIT
    dbg = torch.zeros(96, dtype=torch.int64, device="cuda")
    g.dbg = dbg.data_ptr()
    for _ in range(3):
        _lib.call("tpp_gemm_tc", C.byref(g), _lib.stream_ptr())
    torch.cuda.synchronize()
    t = dbg.cpu().numpy()
    print(f"block_n {bn}: first TMA issue +{t[2] - t[0]}, first MMA wait done +{t[3] - t[0]}, last MMA committed +{t[4] - t[0]}")
    for it in range(12):
        c = t[16 + 4 * it: 20 + 4 * it] - t[0]
        print(f"  k-block {it:2d}: landed +{c[0]:6d}  converted +{c[1] - c[0]:5d}  fenced +{c[2] - c[1]:5d}  arrived +{c[3] - c[2]:5d}"
              f"   MMA saw it +{t[64 + it] - t[0]:6d}")
