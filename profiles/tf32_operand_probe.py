"""Does tcgen05.mma kind::tf32 TRUNCATE or ROUND the fp32 bit patterns it reads from shared memory?  Feeds un-rounded fp32
data as the (single-pass) A operand and compares the product with float64 products of trunc(A) and rna(A)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tpp_b200 import _lib as L  # noqa: E402

M, N, K = 256, 128, 64
torch.manual_seed(0)
a = torch.randn(M, K, device="cuda")
w = torch.randn(N, K, device="cuda")
w_t = ((w.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)            # weights exactly representable
out = torch.zeros(M, N, device="cuda")
g = L.TcGemm()
g.a_hi, g.a_lo, g.lda = a.data_ptr(), a.data_ptr(), K
g.b_hi, g.b_lo, g.ldb = w_t.data_ptr(), w_t.data_ptr(), K
g.M, g.N, g.K, g.precision, g.split_k, g.block_n = M, N, K, 1, 1, 128
g.out, g.ldc = out.data_ptr(), N
L.call("tpp_gemm_tc", L.C.byref(g), L.stream_ptr())
torch.cuda.synchronize()
a_trunc = (a.view(torch.int32) & ~0x1FFF).view(torch.float32)
a_rna = ((a.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)
ref_t = a_trunc.double() @ w_t.double().t()
ref_r = a_rna.double() @ w_t.double().t()
ref_x = a.double() @ w_t.double().t()
for name, ref in (("trunc(A)", ref_t), ("rna(A)", ref_r), ("exact A", ref_x)):
    print(f"max |out - {name} W| / max|ref| = {float((out.double() - ref).abs().max() / ref.abs().max()):.3e}")
