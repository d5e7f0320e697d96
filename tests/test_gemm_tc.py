"""GPU: the tcgen05 tensor-core GEMM (csrc/gemm_tc.cu) against float64 matmul.

Stated tolerances: precision 3 (3xTF32, the parity path): max |err| <= 2e-6 * sum_k |a||b| (fp32-grade);
precision 1 (single TF32 product, fast mode): <= 2e-3 * sum_k |a||b|."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _split(x, ld_out=None, transposed_ld=None):
    from tpp_b200 import _lib
    rows, cols = x.shape
    ld_out = ld_out or (cols + 3) // 4 * 4
    hi = torch.full((rows, ld_out), 7.0, device="cuda")
    lo = torch.full((rows, ld_out), 7.0, device="cuda")
    th = tl = None
    if transposed_ld:
        th = torch.full((cols, transposed_ld), 7.0, device="cuda")
        tl = torch.full((cols, transposed_ld), 7.0, device="cuda")
    _lib.call("tpp_split_tf32", _lib.ptr(x), x.stride(0), rows, cols, _lib.ptr(hi), _lib.ptr(lo), ld_out,
              _lib.ptr(th), _lib.ptr(tl), transposed_ld or 0, _lib.stream_ptr())
    return hi, lo, th, tl


def _operand(x, mn_major):
    """x: logical [rows(M or N), K].  K-major: stored as is (ld = ceil4(K)); MN-major: stored transposed
    [K, ceil32(rows)] via the split kernel's transposed output."""
    rows, K = x.shape
    if not mn_major:
        ld = (K + 3) // 4 * 4
        hi, lo, _, _ = _split(x, ld)
        return hi, lo, ld
    ld = (rows + 31) // 32 * 32
    _, _, th, tl = _split(x, None, ld)
    return th, tl, ld


def _gemm(a, b, precision, flags=0, bias=None, mask=None, split_k=1, block_n=0, want_colsum=False, out=None,
          a_mn=False, b_mn=False, mask_bits_out=None, mask_bits=None):
    from tpp_b200 import _lib
    M, K = a.shape
    N = b.shape[0]
    a_hi, a_lo, lda = _operand(a, a_mn)
    b_hi, b_lo, ldb = _operand(b, b_mn)
    ldc = (N + 3) // 4 * 4
    g = _lib.TcGemm()
    g.a_hi, g.a_lo, g.lda = a_hi.data_ptr(), a_lo.data_ptr(), lda
    g.b_hi, g.b_lo, g.ldb = b_hi.data_ptr(), b_lo.data_ptr(), ldb
    g.a_mn, g.b_mn = int(a_mn), int(b_mn)
    g.M, g.N, g.K, g.precision, g.split_k, g.flags, g.block_n = M, N, K, precision, split_k, flags, block_n
    res = {}
    if out is None:
        res["out"] = torch.full((M, ldc), 5.0, device="cuda")
        res["hi"] = torch.full((M, ldc), 5.0, device="cuda")
        res["lo"] = torch.full((M, ldc), 5.0, device="cuda")
        g.out, g.out_hi, g.out_lo = res["out"].data_ptr(), res["hi"].data_ptr(), res["lo"].data_ptr()
    else:
        res["out"] = out
        g.out = out.data_ptr()
        ldc = out.stride(0)
    g.ldc = ldc
    if want_colsum:
        res["colsum"] = torch.ones(N, device="cuda")
        g.colsum = res["colsum"].data_ptr()
    if bias is not None:
        g.bias = bias.data_ptr()
    if mask is not None:
        g.mask, g.ld_mask = mask.data_ptr(), mask.stride(0)
    if mask_bits_out is not None:
        g.mask_bits_out = mask_bits_out.data_ptr()
    if mask_bits is not None:
        g.mask_bits = mask_bits.data_ptr()
    _lib.call("tpp_gemm_tc", C.byref(g), _lib.stream_ptr())
    torch.cuda.synchronize()
    return res


def _ref(a, b):
    a64, b64 = a.double().cpu(), b.double().cpu()
    return a64 @ b64.t(), a64.abs() @ b64.abs().t()


@pytest.mark.parametrize("rows,cols", [(64, 32), (100, 9), (5, 588), (256, 256)])
def test_split_pairs_are_exact(rows, cols):
    x = torch.randn(rows, cols, device="cuda") * 3
    ld, ldt = (cols + 3) // 4 * 4 + 4, (rows + 3) // 4 * 4
    hi, lo, th, tl = _split(x, ld, ldt)
    assert torch.equal(hi[:, :cols] + lo[:, :cols], x)                 # exact decomposition
    assert (hi[:, cols:] == 0).all() and (lo[:, cols:] == 0).all()      # padding is zero
    assert torch.equal(hi.view(torch.int32) & 0x1FFF, torch.zeros_like(hi, dtype=torch.int32))   # tf32-representable
    assert (lo.abs() <= hi.abs() * 2.0 ** -11 + 1e-38).all()
    assert torch.equal(th[:, :rows], hi[:, :cols].t()) and torch.equal(tl[:, :rows], lo[:, :cols].t())
    assert (th[:, rows:] == 0).all()


@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (128, 64, 256), (256, 256, 256), (8192, 256, 588), (4096, 64, 256),
                                   (128, 5, 64), (384, 256, 12), (130, 70, 100), (128, 256, 8192)])
@pytest.mark.parametrize("precision", [3, 1])
def test_plain_product(M, N, K, precision):
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randn(M, K, device="cuda", generator=g)
    b = torch.randn(N, K, device="cuda", generator=g)
    res = _gemm(a, b, precision)
    want, scale = _ref(a, b)
    err = (res["out"][:, :N].double().cpu() - want).abs() / scale
    # fp32 accumulation in TMEM over K terms: the bound grows with K (an unsplit K=8192 contraction measures 5e-6);
    # the engine splits long contractions (weight gradients) across CTAs, which also shortens each accumulation.
    tol = (2e-6 if precision == 3 else 2e-3) * max(1.0, K / 2048)
    assert err.max() < tol, f"max scaled err {err.max():.3e}"
    assert torch.equal(res["hi"][:, :N] + res["lo"][:, :N], res["out"][:, :N])


@pytest.mark.parametrize("block_n", [16, 64, 128, 256])
def test_block_n_variants(block_n):
    a = torch.randn(256, 96, device="cuda")
    b = torch.randn(block_n * 2 - 8, 96, device="cuda")
    res = _gemm(a, b, 3, block_n=block_n)
    want, scale = _ref(a, b)
    assert ((res["out"][:, :b.shape[0]].double().cpu() - want).abs() / scale).max() < 2e-6


def test_epilogue_bias_relu_mask_and_transposed_outputs():
    M, N, K = 384, 256, 200
    a, b = torch.randn(M, K, device="cuda"), torch.randn(N, K, device="cuda")
    bias = torch.randn(N, device="cuda")
    mask = torch.randn(M, N, device="cuda")
    want, scale = _ref(a, b)
    res = _gemm(a, b, 3, flags=1 | 2, bias=bias)                        # bias + relu (forward layer)
    w = torch.relu(want + bias.double().cpu())
    assert ((res["out"].double().cpu() - w).abs() / (scale + 1)).max() < 2e-6
    res = _gemm(a, b, 3, flags=4, mask=mask, want_colsum=True)            # relu-mask + column sums (data gradient)
    w = want * (mask.cpu() > 0)
    assert ((res["out"].double().cpu() - w).abs() / scale).max() < 2e-6
    cs = res["colsum"].double().cpu() - 1.0                               # accumulated on top of the initial ones
    np.testing.assert_allclose(cs.numpy(), w.sum(0).numpy(), rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("a_mn,b_mn", [(True, False), (False, True), (True, True)])
@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (256, 256, 100), (4096, 256, 9), (8192, 588, 256), (256, 608, 1000),
                                   (130, 70, 50)])
def test_mn_major_operands(a_mn, b_mn, M, N, K):
    """Rollout layer 1 (feature-major obs: A MN-major), data gradient (W as MN-major B), weight gradient (both)."""
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    a = torch.randn(M, K, device="cuda", generator=g)
    b = torch.randn(N, K, device="cuda", generator=g)
    res = _gemm(a, b, 3, a_mn=a_mn, b_mn=b_mn)
    want, scale = _ref(a, b)
    err = (res["out"][:, :N].double().cpu() - want).abs() / scale
    assert err.max() < 2e-6, f"max scaled err {err.max():.3e}"


@pytest.mark.parametrize("split_k", [1, 7, 64])
@pytest.mark.parametrize("mn", [False, True])
def test_split_k_atomic_accumulation(split_k, mn):
    """Weight-gradient shape: small output, long contraction over the minibatch, accumulated on top of `out`
    (mn=True: both operands MN-major, i.e. read straight from the row-major dZ [mb][out] and X [mb][in])."""
    a, b = torch.randn(256, 8192, device="cuda"), torch.randn(588, 8192, device="cuda")
    base = torch.randn(256, 588, device="cuda")
    out = base.clone()
    _gemm(a, b, 3, flags=8, split_k=split_k, out=out, a_mn=mn, b_mn=mn)
    want, scale = _ref(a, b)
    err = (out.double().cpu() - (base.double().cpu() + want)).abs() / scale
    assert err.max() < (2e-6 if split_k > 1 else 8e-6)       # unsplit: 8192-term fp32 accumulation


@pytest.mark.parametrize("which", ["a", "b"])
@pytest.mark.parametrize("M,N,K,mn", [(256, 128, 588, False), (130, 70, 100, False), (64, 588, 512, True)])
def test_exact_operand_two_pass_mode(which, M, N, K, mn):
    """TPP_TC_A_EXACT / TPP_TC_B_EXACT: the operand holds integer pixel values 0..255 (exact in TF32); its lo half is
    neither passed nor loaded and the result still matches float64 to fp32 grade.  MN-major case = the first layer's
    weight gradient (contraction over samples, alpha = 1/255 folded into the atomic accumulation)."""
    from tpp_b200 import _lib
    g0 = torch.Generator(device="cuda").manual_seed(M * 3 + N + K)
    a = torch.randn(M, K, device="cuda", generator=g0)
    b = torch.randn(N, K, device="cuda", generator=g0)
    pix = lambda t: torch.randint(0, 256, t.shape, device="cuda", generator=g0).float()
    if which == "a":
        a = pix(a)
    else:
        b = pix(b)
    a_hi, a_lo, lda = _operand(a, mn)
    b_hi, b_lo, ldb = _operand(b, mn)
    assert (a_lo if which == "a" else b_lo).abs().max() == 0          # really exact
    ldc = (N + 3) // 4 * 4
    out = torch.zeros(M, ldc, device="cuda")
    g = _lib.TcGemm()
    g.a_hi, g.lda, g.b_hi, g.ldb = a_hi.data_ptr(), lda, b_hi.data_ptr(), ldb
    if which == "a":
        g.b_lo = b_lo.data_ptr()
    else:
        g.a_lo = a_lo.data_ptr()
    g.a_mn = g.b_mn = int(mn)
    g.M, g.N, g.K = M, N, K
    g.precision = 3 | (_lib.TC_A_EXACT if which == "a" else _lib.TC_B_EXACT)
    g.split_k, g.flags, g.out, g.ldc = (4 if mn else 1), (_lib.EPI_ACCUM if mn else 0), out.data_ptr(), ldc
    g.alpha = 1.0 / 255.0 if mn else 0.0
    _lib.call("tpp_gemm_tc", C.byref(g), _lib.stream_ptr())
    torch.cuda.synchronize()
    want, scale = _ref(a, b)
    if mn:
        want, scale = want / 255.0, scale / 255.0
    err = (out[:, :N].double().cpu() - want).abs() / scale
    assert err.max() < 2e-6, f"max scaled err {err.max():.3e}"


# ---- 256 x 256 tile on a CTA pair (cta_group::2, block_n = 512) -------------------------------------------------
@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, True)])
@pytest.mark.parametrize("M,N,K", [(256, 256, 32), (512, 256, 256), (4096, 256, 588), (1000, 256, 200), (256, 588, 96),
                                   (300, 500, 72)])
@pytest.mark.parametrize("precision", [3, 1])
@pytest.mark.parametrize("block_n", [512, 513, 514])
def test_cta_pair_tile(a_mn, b_mn, M, N, K, precision, block_n):
    """Forward (both K-major), data gradient (B MN-major) and weight-gradient (both MN-major) operand forms; ragged M / N
    edges (rows and columns beyond the matrix come from TMA zero fill and are never stored)."""
    g = torch.Generator(device="cuda").manual_seed(M * 5 + N * 3 + K + 17)
    a = torch.randn(M, K, device="cuda", generator=g)
    b = torch.randn(N, K, device="cuda", generator=g)
    res = _gemm(a, b, precision, a_mn=a_mn, b_mn=b_mn, block_n=block_n)
    want, scale = _ref(a, b)
    err = (res["out"][:, :N].double().cpu() - want).abs() / scale
    assert err.max() < (2e-6 if precision == 3 else 2e-3), f"max scaled err {err.max():.3e}"
    assert torch.equal(res["hi"][:, :N] + res["lo"][:, :N], res["out"][:, :N])


@pytest.mark.parametrize("tile", [513, 514])
@pytest.mark.parametrize("M,N,K,b_mn", [(256 * 200, 256, 96, False), (256 * 163 + 40, 256, 64, True), (256 * 90, 500, 40, False)])
def test_persistent_cta_pairs_loop_over_work_items(M, N, K, b_mn, tile):
    """block_n = 513: 74 CTA pairs loop over more 256 x 256 work items than there are pairs (two TMEM accumulators,
    the epilogue of one item overlapping the loads / MMAs of the next), with the fused epilogue pieces."""
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randn(M, K, device="cuda", generator=g)
    b = torch.randn(N, K, device="cuda", generator=g)
    bias, mask = torch.randn(N, device="cuda", generator=g), torch.randn(M, N, device="cuda", generator=g)
    want, scale = _ref(a, b)
    res = _gemm(a, b, 3, flags=1 | 2, bias=bias, b_mn=b_mn, block_n=tile)
    w = torch.relu(want + bias.double().cpu())
    assert ((res["out"][:, :N].double().cpu() - w).abs() / (scale + 1)).max() < 2e-6
    assert torch.equal(res["hi"][:, :N] + res["lo"][:, :N], res["out"][:, :N])
    res = _gemm(a, b, 3, flags=4, mask=mask, want_colsum=True, b_mn=b_mn, block_n=tile)
    w = want * (mask.cpu() > 0)
    assert ((res["out"][:, :N].double().cpu() - w).abs() / scale).max() < 2e-6
    np.testing.assert_allclose(res["colsum"].double().cpu().numpy() - 1.0, w.sum(0).numpy(), rtol=1e-4, atol=2e-2)


@pytest.mark.parametrize("block_n", [512, 513, 514])
def test_cta_pair_weight_gradient_many_splits(block_n):
    a2, b2 = torch.randn(256, 131072 // 4, device="cuda"), torch.randn(588, 131072 // 4, device="cuda")
    out = torch.ones(256, 588, device="cuda")
    _gemm(a2, b2, 3, flags=8, split_k=64, a_mn=True, b_mn=True, out=out, block_n=block_n)
    want2, scale2 = _ref(a2, b2)
    assert ((out.double().cpu() - 1.0 - want2).abs() / scale2).max() < 2e-6


def test_cta_pair_epilogues_and_split_k():
    M, N, K = 1024, 256, 200
    a, b = torch.randn(M, K, device="cuda"), torch.randn(N, K, device="cuda")
    bias, mask = torch.randn(N, device="cuda"), torch.randn(M, N, device="cuda")
    want, scale = _ref(a, b)
    res = _gemm(a, b, 3, flags=1 | 2, bias=bias, block_n=512)                       # bias + relu
    w = torch.relu(want + bias.double().cpu())
    assert ((res["out"].double().cpu() - w).abs() / (scale + 1)).max() < 2e-6
    res = _gemm(a, b, 3, flags=4, mask=mask, want_colsum=True, b_mn=True, block_n=512)   # data-gradient form
    w = want * (mask.cpu() > 0)
    assert ((res["out"].double().cpu() - w).abs() / scale).max() < 2e-6
    np.testing.assert_allclose(res["colsum"].double().cpu().numpy() - 1.0, w.sum(0).numpy(), rtol=1e-4, atol=1e-3)
    # weight-gradient form: long contraction split over CTAs pairs, atomically accumulated on top of existing values
    a2, b2 = torch.randn(256, 8192, device="cuda"), torch.randn(588, 8192, device="cuda")
    out = torch.ones(256, 588, device="cuda")
    _gemm(a2, b2, 3, flags=8, split_k=16, a_mn=True, b_mn=True, out=out, block_n=512)
    want2, scale2 = _ref(a2, b2)
    assert ((out.double().cpu() - 1.0 - want2).abs() / scale2).max() < 2e-6


@pytest.mark.parametrize("M,N,K", [(256 * 180 + 100, 64, 256), (512, 64, 96), (256 * 80, 40, 64)])
def test_persistent_cta_pairs_narrow_output(M, N, K):
    """block_n = 65: 256 x 64 tiles on persistent CTA pairs (the 256 -> 64 layer of a fused accumulation window)."""
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randn(M, K, device="cuda", generator=g)
    b = torch.randn(N, K, device="cuda", generator=g)
    bias = torch.randn(N, device="cuda", generator=g)
    want, scale = _ref(a, b)
    res = _gemm(a, b, 3, flags=1, bias=bias, block_n=65)
    w = want + bias.double().cpu()
    assert ((res["out"][:, :N].double().cpu() - w).abs() / (scale + 1)).max() < 2e-6
    assert torch.equal(res["hi"][:, :N] + res["lo"][:, :N], res["out"][:, :N])


# ---- operands split on chip (TPP_TC_A_SPLIT / TPP_TC_B_SPLIT) ----------------------------------------------------
def _plain_operand(x, mn_major):
    """The plain fp32 array a producer would leave in HBM (no hi / lo pair)."""
    rows, K = x.shape
    if not mn_major:
        ld = (K + 3) // 4 * 4
        buf = torch.zeros(rows, ld, device="cuda")
        buf[:, :K] = x
        return buf, ld
    ld = (rows + 31) // 32 * 32
    buf = torch.zeros(K, ld, device="cuda")
    buf[:, :rows] = x.t()
    return buf, ld


@pytest.mark.parametrize("which", ["a", "b", "ab"])
@pytest.mark.parametrize("block_n", [64, 128, 256, 512, 513, 65])
@pytest.mark.parametrize("M,N,K,a_mn,b_mn", [(512, 256, 256, False, False), (1000, 256, 200, False, True),
                                             (256, 300, 1024, True, True), (256 * 160 + 40, 64, 96, False, False)])
def test_operands_split_on_chip(which, block_n, M, N, K, a_mn, b_mn):
    """The operand is handed over as ONE plain fp32 array; four extra warps form its lo half in shared memory.  Same
    fp32-grade bound as the pair path, on every wide tile, K-major and MN-major, with the other operand a pair."""
    from tpp_b200 import _lib
    if block_n == 65 and N > 64:
        pytest.skip("256 x 64 tile: N <= 64")
    atomic = a_mn and b_mn                      # the weight-gradient form: split-k, atomic accumulation
    g0 = torch.Generator(device="cuda").manual_seed(M + 7 * N + K)
    a = torch.randn(M, K, device="cuda", generator=g0)
    b = torch.randn(N, K, device="cuda", generator=g0) * 0.05
    g = _lib.TcGemm()
    prec = 3
    if "a" in which:
        buf_a, lda = _plain_operand(a, a_mn)
        g.a_hi, prec = buf_a.data_ptr(), prec | _lib.TC_A_SPLIT
    else:
        a_hi, a_lo, lda = _operand(a, a_mn)
        g.a_hi, g.a_lo = a_hi.data_ptr(), a_lo.data_ptr()
    if "b" in which:
        buf_b, ldb = _plain_operand(b, b_mn)
        g.b_hi, prec = buf_b.data_ptr(), prec | _lib.TC_B_SPLIT
    else:
        b_hi, b_lo, ldb = _operand(b, b_mn)
        g.b_hi, g.b_lo = b_hi.data_ptr(), b_lo.data_ptr()
    ldc = (N + 3) // 4 * 4
    out = torch.zeros(M, ldc, device="cuda")
    g.lda, g.ldb, g.a_mn, g.b_mn = lda, ldb, int(a_mn), int(b_mn)
    g.M, g.N, g.K, g.precision, g.block_n = M, N, K, prec, block_n
    g.split_k, g.flags, g.out, g.ldc = (5 if atomic else 1), (_lib.EPI_ACCUM if atomic else 0), out.data_ptr(), ldc
    _lib.call("tpp_gemm_tc", C.byref(g), _lib.stream_ptr())
    torch.cuda.synchronize()
    want, scale = _ref(a, b)
    err = (out[:, :N].double().cpu() - want).abs() / scale
    assert err.max() < 2e-6, f"max scaled err {err.max():.3e}"


def test_split_on_chip_refused_on_narrow_tiles():
    from tpp_b200 import _lib
    a = torch.randn(128, 64, device="cuda")
    b = torch.randn(16, 64, device="cuda")
    out = torch.zeros(128, 16, device="cuda")
    g = _lib.TcGemm()
    g.a_hi, g.lda, g.b_hi, g.ldb = a.data_ptr(), 64, b.data_ptr(), 64
    g.M, g.N, g.K, g.precision, g.block_n, g.out, g.ldc = 128, 16, 64, 3 | _lib.TC_A_SPLIT | _lib.TC_B_SPLIT, 16, \
        out.data_ptr(), 16
    with pytest.raises(Exception):
        _lib.call("tpp_gemm_tc", C.byref(g), _lib.stream_ptr())


@pytest.mark.parametrize("block_n", [64, 128, 256, 512, 513, 514, 65])
def test_one_bit_relu_masks(block_n):
    """mask_bits_out / mask_bits: the forward epilogue leaves (relu output > 0) as one bit per element, the data
    gradient's epilogue masks with those bits -- same result, bit for bit, as TPP_EPI_MASK on the fp32 activation."""
    M, N, K = 256 * 5 + 64, (64 if block_n == 65 else 256), 96
    g0 = torch.Generator(device="cuda").manual_seed(block_n)
    a = torch.randn(M, K, device="cuda", generator=g0)
    b = torch.randn(N, K, device="cuda", generator=g0) * 0.1
    bias = torch.randn(N, device="cuda", generator=g0) * 0.1
    bits = torch.full((M // 32 * (N // 32) * 32,), -1, dtype=torch.int32, device="cuda")
    fwd = _gemm(a, b, 3, flags=1 | 2, bias=bias, block_n=block_n, mask_bits_out=bits)
    H = fwd["out"][:, :N].contiguous()
    assert 0.3 < (H > 0).float().mean() < 0.7
    a2 = torch.randn(M, 64, device="cuda", generator=g0)
    b2 = torch.randn(N, 64, device="cuda", generator=g0)
    ref = _gemm(a2, b2, 3, flags=4, mask=H, want_colsum=True, block_n=block_n)
    got = _gemm(a2, b2, 3, want_colsum=True, block_n=block_n, mask_bits=bits)
    for k in ("out", "hi", "lo"):
        assert torch.equal(ref[k], got[k]), k
    torch.testing.assert_close(ref["colsum"], got["colsum"], rtol=1e-5, atol=1e-4)     # atomics: order varies per run
    assert ((got["out"][:, :N] != 0) <= (H > 0)).all()


def test_one_bit_masks_refused_where_the_fast_path_does_not_run():
    a = torch.randn(100, 64, device="cuda")        # rows not a multiple of 32
    b = torch.randn(256, 64, device="cuda")
    bits = torch.zeros(4 * 8 * 32, dtype=torch.int32, device="cuda")
    with pytest.raises(Exception):
        _gemm(a, b, 3, flags=2, block_n=128, mask_bits_out=bits)
