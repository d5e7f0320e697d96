"""GPU: the IMPALA-CNN building blocks (csrc/conv.cu + tpp_gemm_tc) against torch fp32 / fp64 references of the same
ops (reference common/model.py:134-208): 3x3 convolution forward, data gradient and weight gradient as im2col +
tensor-core GEMM, max-pool 3x3/2 forward and backward, narrow column sums, and the whole ImpalaEngineTC forward /
backward against torch autograd on the same flat parameters.  Tolerances: 3xTF32 is fp32-grade -> 2e-5 relative to
the output scale for single layers, 1e-4 for gradients that contract over B*H*W >= 10^4 terms."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _lib():
    from tpp_b200 import _lib
    return _lib


def _probe():
    """tests/native/libtpp_probe.so: the test-only TMA im2col probe (not part of the product library)."""
    import ctypes
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "native", "libtpp_probe.so")
    if not os.path.exists(path):
        pytest.skip("tests/native/libtpp_probe.so not built (make -C tests/native)")
    lib = ctypes.CDLL(path)
    lib.tpp_debug_tma_im2col.argtypes = [ctypes.c_void_p] + [ctypes.c_int32] * 12 + [ctypes.c_void_p] * 2
    lib.tpp_debug_tma_im2col.restype = ctypes.c_int
    return lib


def _pair(x):
    hi = ((x.contiguous().view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)
    return hi, x - hi


def _close(a, b, tol):
    a, b = a.double().cpu(), b.double().cpu()
    scale = b.abs().max().item() + 1e-30
    err = (a - b).abs().max().item()
    assert err <= tol * scale, f"max err {err:.3e} vs scale {scale:.3e} (tol {tol})"


@pytest.mark.parametrize("B,H,W,C,relu,u8", [(2, 8, 8, 16, 0, 0), (3, 7, 5, 32, 1, 0), (2, 14, 14, 3, 0, 0),
                                             (2, 9, 6, 3, 0, 1)])
def test_im2col_matches_unfold(B, H, W, C, relu, u8):
    L = _lib()
    torch.manual_seed(0)
    if u8:
        x = torch.randint(0, 256, (B, H, W, C), dtype=torch.uint8, device="cuda")
        xf, scale = x.float() / 255.0, 1.0 / 255.0
    else:
        x = torch.randn(B, H, W, C, device="cuda")
        xf, scale = x, 1.0
    Kp = (9 * C + 31) // 32 * 32
    hi = torch.full((B * H * W, Kp), 7.0, device="cuda")
    lo = torch.full((B * H * W, Kp), 7.0, device="cuda")
    L.call("tpp_im2col3x3", L.ptr(x), u8, B, H, W, C, H * W * C, W * C, C, 1, relu, scale, L.ptr(hi), L.ptr(lo), Kp,
           L.stream_ptr())
    ref_in = (F.relu(xf) if relu else xf).permute(0, 3, 1, 2)
    # unfold: [B, C*9, HW] with index c*9 + tap -> reorder to tap*C + c
    ref = F.unfold(ref_in, 3, padding=1).view(B, C, 9, H * W).permute(0, 3, 2, 1).reshape(B * H * W, 9 * C)
    got = hi + lo
    if u8:
        torch.testing.assert_close(got[:, :9 * C], ref, rtol=1e-6, atol=1e-7)   # x*(1/255) vs x/255
    else:
        assert torch.equal(got[:, :9 * C], ref)
    assert torch.equal(hi[:, :9 * C], _pair(got[:, :9 * C].contiguous())[0])
    assert (hi[:, 9 * C:] == 0).all() and (lo[:, 9 * C:] == 0).all()


def test_im2col_reads_nchw_rows_through_strides():
    L = _lib()
    B, C, H, W = 3, 3, 10, 12
    ld = C * H * W + 8
    x = torch.randn(B, ld, device="cuda")
    Kp = 32
    hi = torch.zeros(B * H * W, Kp, device="cuda")
    lo = torch.zeros_like(hi)
    L.call("tpp_im2col3x3", L.ptr(x), 0, B, H, W, C, ld, W, 1, H * W, 0, 1.0, L.ptr(hi), L.ptr(lo), Kp, L.stream_ptr())
    img = x[:, :C * H * W].reshape(B, C, H, W)
    ref = F.unfold(img, 3, padding=1).view(B, C, 9, H * W).permute(0, 3, 2, 1).reshape(B * H * W, 9 * C)
    assert torch.equal((hi + lo)[:, :27], ref)


def _conv_gemm(L, col, w_pair, rows, cout, Kp, **kw):
    g = L.TcGemm()
    g.a_hi, g.a_lo, g.lda = col[0].data_ptr(), col[1].data_ptr(), Kp
    g.b_hi, g.b_lo, g.ldb = w_pair[0].data_ptr(), w_pair[1].data_ptr(), Kp
    g.M, g.N, g.K, g.precision, g.split_k = rows, cout, Kp, 3, 1
    for k, v in kw.items():
        setattr(g, k, v)
    L.call("tpp_gemm_tc", L.C.byref(g), L.stream_ptr())


@pytest.mark.parametrize("B,H,W,cin,cout", [(4, 16, 16, 16, 16), (2, 32, 32, 16, 32), (3, 8, 8, 32, 32),
                                            (2, 64, 64, 3, 16), (5, 7, 7, 32, 32)])
def test_conv3x3_forward_residual_relu(B, H, W, cin, cout):
    L = _lib()
    torch.manual_seed(1)
    x = torch.randn(B, H, W, cin, device="cuda")
    w = torch.randn(cout, cin, 3, 3, device="cuda") * 0.1
    bias = torch.randn(cout, device="cuda")
    skip = torch.randn(B * H * W, cout, device="cuda")
    Kp = (9 * cin + 31) // 32 * 32
    rows = B * H * W
    col = (torch.zeros(rows, Kp, device="cuda"), torch.zeros(rows, Kp, device="cuda"))
    L.call("tpp_im2col3x3", L.ptr(x), 0, B, H, W, cin, H * W * cin, W * cin, cin, 1, 1, 1.0, L.ptr(col[0]),
           L.ptr(col[1]), Kp, L.stream_ptr())
    wf = torch.zeros(cout, Kp, device="cuda")
    wf[:, :9 * cin] = w.permute(0, 2, 3, 1).reshape(cout, 9 * cin)
    out = torch.zeros(rows, cout, device="cuda")
    oh, ol = torch.zeros_like(out), torch.zeros_like(out)
    _conv_gemm(L, col, _pair(wf), rows, cout, Kp, flags=L.EPI_BIAS | L.EPI_ADD | L.EPI_RELU_OUT,
               bias=bias.data_ptr(), addend=skip.data_ptr(), ld_add=cout, out=out.data_ptr(), out_hi=oh.data_ptr(),
               out_lo=ol.data_ptr(), ldc=cout)
    ref = F.conv2d(F.relu(x.double()).permute(0, 3, 1, 2), w.double(), bias.double(), padding=1)
    ref = F.relu(ref.permute(0, 2, 3, 1).reshape(rows, cout) + skip.double())
    _close(out, ref, 2e-5)
    assert torch.equal(oh + ol, out)


@pytest.mark.parametrize("B,H,W,cin,cout", [(4, 16, 16, 16, 16), (2, 32, 32, 16, 32), (3, 8, 8, 32, 32),
                                            (2, 14, 14, 3, 16)])
def test_conv3x3_weight_gradient(B, H, W, cin, cout):
    """dWf = dY^T col(X): both operands MN-major, 16-wide dY exercises the narrow (ld < 32) tensor map."""
    L = _lib()
    torch.manual_seed(2)
    x = torch.randn(B, H, W, cin, device="cuda")
    dy = torch.randn(B * H * W, cout, device="cuda")
    Kp = (9 * cin + 31) // 32 * 32
    rows = B * H * W
    col = (torch.zeros(rows, Kp, device="cuda"), torch.zeros(rows, Kp, device="cuda"))
    L.call("tpp_im2col3x3", L.ptr(x), 0, B, H, W, cin, H * W * cin, W * cin, cin, 1, 0, 1.0, L.ptr(col[0]),
           L.ptr(col[1]), Kp, L.stream_ptr())
    dyp = _pair(dy)
    for bn, split in ((128, 1), (256, 7), (128, 64)):
        gw = torch.zeros(cout, Kp, device="cuda")
        g = L.TcGemm()
        g.a_hi, g.a_lo, g.lda = dyp[0].data_ptr(), dyp[1].data_ptr(), cout
        g.b_hi, g.b_lo, g.ldb = col[0].data_ptr(), col[1].data_ptr(), Kp
        g.M, g.N, g.K, g.precision, g.split_k, g.a_mn, g.b_mn = cout, 9 * cin, rows, 3, split, 1, 1
        g.flags, g.out, g.ldc, g.block_n = L.EPI_ACCUM, gw.data_ptr(), Kp, bn
        L.call("tpp_gemm_tc", L.C.byref(g), L.stream_ptr())
        xd = x.double().permute(0, 3, 1, 2).requires_grad_(True)
        wd = torch.zeros(cout, cin, 3, 3, dtype=torch.float64, device="cuda", requires_grad=True)
        y = F.conv2d(xd, wd, padding=1)
        y.backward(dy.double().view(B, H, W, cout).permute(0, 3, 1, 2))
        ref = wd.grad.permute(0, 2, 3, 1).reshape(cout, 9 * cin)
        _close(gw[:, :9 * cin], ref, 1e-4)
        assert (gw[:, 9 * cin:] == 0).all()


@pytest.mark.parametrize("B,H,W,cin,cout", [(4, 16, 16, 16, 16), (2, 32, 32, 16, 32), (3, 9, 9, 32, 32)])
def test_conv3x3_data_gradient_mask_skip_colsum(B, H, W, cin, cout):
    L = _lib()
    torch.manual_seed(3)
    xpre = torch.randn(B * H * W, cin, device="cuda")          # pre-ReLU input of the conv (mask)
    w = torch.randn(cout, cin, 3, 3, device="cuda") * 0.1
    dy = torch.randn(B, H, W, cout, device="cuda")
    skip = torch.randn(B * H * W, cin, device="cuda")
    Kd = (9 * cout + 31) // 32 * 32
    rows = B * H * W
    col = (torch.zeros(rows, Kd, device="cuda"), torch.zeros(rows, Kd, device="cuda"))
    L.call("tpp_im2col3x3", L.ptr(dy), 0, B, H, W, cout, H * W * cout, W * cout, cout, 1, 0, 1.0, L.ptr(col[0]),
           L.ptr(col[1]), Kd, L.stream_ptr())
    wdm = torch.zeros(cin, Kd, device="cuda")
    wdm[:, :9 * cout] = w.flip(2, 3).permute(1, 2, 3, 0).reshape(cin, 9 * cout)
    dx = torch.zeros(rows, cin, device="cuda")
    dh, dl = torch.zeros_like(dx), torch.zeros_like(dx)
    cs = torch.zeros(cin, device="cuda")
    _conv_gemm(L, col, _pair(wdm), rows, cin, Kd, flags=L.EPI_MASK | L.EPI_ADD, mask=xpre.data_ptr(), ld_mask=cin,
               addend=skip.data_ptr(), ld_add=cin, out=dx.data_ptr(), out_hi=dh.data_ptr(), out_lo=dl.data_ptr(),
               ldc=cin, colsum=cs.data_ptr())
    xd = xpre.double().view(B, H, W, cin).permute(0, 3, 1, 2).requires_grad_(True)
    y = F.conv2d(F.relu(xd), w.double(), padding=1)
    y.backward(dy.double().permute(0, 3, 1, 2))
    ref = xd.grad.permute(0, 2, 3, 1).reshape(rows, cin) + skip.double()
    _close(dx, ref, 2e-5)
    _close(cs, ref.sum(0), 1e-4)
    assert torch.equal(dh + dl, dx)


@pytest.mark.parametrize("B,H,W,C", [(3, 64, 64, 16), (2, 14, 14, 16), (2, 7, 7, 32), (4, 5, 9, 32), (1, 2, 2, 16)])
def test_maxpool_forward_backward(B, H, W, C):
    L = _lib()
    torch.manual_seed(4)
    x = torch.randn(B, H, W, C, device="cuda")
    x[0, :, :, 0] = 0.5                                           # ties: first maximum wins, like torch
    Ho, Wo = (H + 1) // 2, (W + 1) // 2
    y = torch.zeros(B, Ho, Wo, C, device="cuda")
    arg = torch.zeros(B, Ho, Wo, C, dtype=torch.uint8, device="cuda")
    rh, rl = torch.zeros_like(y), torch.zeros_like(y)
    L.call("tpp_maxpool3x3s2_fwd", L.ptr(x), B, H, W, C, L.ptr(y), L.ptr(arg), L.ptr(rh), L.ptr(rl), L.stream_ptr())
    assert torch.equal(rh + rl, y.clamp_min(0))
    xt = x.permute(0, 3, 1, 2).clone().requires_grad_(True)
    ref = F.max_pool2d(xt, 3, 2, 1)
    assert torch.equal(y.permute(0, 3, 1, 2), ref)
    dy = torch.randn(B, Ho, Wo, C, device="cuda")
    ref.backward(dy.permute(0, 3, 1, 2))
    dx = torch.zeros_like(x)
    dh, dl = torch.zeros_like(x), torch.zeros_like(x)
    L.call("tpp_maxpool3x3s2_bwd", L.ptr(dy), L.ptr(arg), B, H, W, C, L.ptr(dx), L.ptr(dh), L.ptr(dl), L.stream_ptr())
    torch.testing.assert_close(dx.permute(0, 3, 1, 2), xt.grad, rtol=1e-6, atol=1e-6)
    assert torch.equal(dh + dl, dx)


@pytest.mark.parametrize("M,C", [(1000, 16), (4097, 32), (64 * 64 * 8, 16), (3, 64)])
def test_colsum_narrow(M, C):
    L = _lib()
    x = torch.randn(M, C, device="cuda")
    out = torch.ones(C, device="cuda")
    L.call("tpp_colsum_narrow", L.ptr(x), M, C, L.ptr(out), L.stream_ptr())
    _close(out, 1.0 + x.double().sum(0), 1e-5)


@pytest.mark.parametrize("hw,B,A", [((64, 64), 6, 15), ((14, 14), 9, 4)])
def test_impala_engine_forward_backward_vs_autograd(hw, B, A):
    """Whole embedder + heads: head outputs and every parameter gradient against torch autograd of a float64 copy of
    the same parameters (the fp32 cuDNN path's own error against float64 is printed for scale: the hand-written
    path must be at least as close)."""
    import copy
    from tpp_b200.common.engine import ImpalaEngineTC
    from tpp_b200.common.model import ImpalaModel
    from tpp_b200.common.policy import CategoricalPolicy
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(5)
    pol = CategoricalPolicy(ImpalaModel(3, input_hw=hw), False, A).to("cuda")
    pol64 = copy.deepcopy(pol).double()
    pol.flatten_()
    eng = ImpalaEngineTC(pol, A, (3, *hw))
    ld = 3 * hw[0] * hw[1]
    x = torch.rand(B, ld, device="cuda")
    head = eng.forward(x, B, train=True)
    feat64, _, fs64, _ = pol64.embedder.forward_with_attn_indices(x.double().view(B, 3, *hw))
    ref_head = torch.cat((pol64.fc_policy(feat64), pol64.fc_value(feat64)), 1)
    feat32, _, _, _ = pol.embedder.forward_with_attn_indices(x.view(B, 3, *hw))
    lib_head = torch.cat((pol.fc_policy(feat32), pol.fc_value(feat32)), 1)
    scale = ref_head.abs().max().item()
    err_tc = (head[:, :A + 1].double() - ref_head).abs().max().item() / scale
    err_lib = (lib_head.double() - ref_head).abs().max().item() / scale
    print(f"head rel err vs float64: hand-written {err_tc:.2e}, cuDNN fp32 {err_lib:.2e}")
    # 17 layers deep; the tensor core adds into its fp32 accumulator with truncation, which costs ~1e-5 over cuDNN
    assert err_tc <= max(5e-5, 2.0 * err_lib)
    np.testing.assert_allclose(eng.last_fs.item(), fs64.item(), rtol=1e-5)
    dhead = torch.zeros(B, eng.ld_head, device="cuda")
    dhead[:, :A + 1] = torch.randn(B, A + 1, device="cuda")
    ref_head.backward(dhead[:, :A + 1].double())
    ref = {n: p.grad for n, p in pol64.named_parameters()}
    pol.flat_grad.zero_()
    lib_head.backward(dhead[:, :A + 1])
    lib_grad = pol.flat_grad.clone()
    pol.flat_grad.zero_()
    eng.backward(dhead, B)
    worst_tc = worst_lib = 0.0
    for name, (off, shape) in pol.layout.items():
        n = int(np.prod(shape))
        b = ref[name].reshape(-1)
        scale = b.abs().max().item() + 1e-12
        e_tc = (pol.flat_grad[off:off + n].double() - b).abs().max().item() / scale
        e_lib = (lib_grad[off:off + n].double() - b).abs().max().item() / scale
        worst_tc, worst_lib = max(worst_tc, e_tc), max(worst_lib, e_lib)
        assert e_tc <= max(1e-4, 3.0 * e_lib), f"{name}: rel err {e_tc:.3e} (cuDNN fp32: {e_lib:.3e})"
    print(f"gradient rel err vs float64: hand-written {worst_tc:.2e}, cuDNN fp32 {worst_lib:.2e}")


def _probe_expected(x, cpp, pixels, w, h, n, ow, oh):
    B, H, W, C = x.shape
    out = torch.zeros(pixels, cpp)
    xc = x.cpu()
    q, p, b = w + 1, h + 1, n
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


@pytest.mark.parametrize("B,H,W,C,cpp,base,tap", [
    (3, 8, 8, 32, 32, (-1, -1, 0), (0, 0)), (3, 8, 8, 32, 32, (-1, -1, 0), (2, 2)), (3, 8, 8, 32, 32, (2, 4, 1), (1, 1)),
    (3, 8, 8, 32, 32, (-1, 5, 2), (1, 1)), (3, 8, 8, 16, 32, (-1, -1, 0), (1, 1)), (3, 8, 8, 4, 32, (-1, -1, 0), (2, 0)),
    (4, 7, 5, 16, 32, (1, 2, 0), (2, 1))])
def test_tma_im2col_conventions(B, H, W, C, cpp, base, tap):
    """Pins what the implicit-GEMM path assumes about TMA im2col loads: bounding box corners (-1, -1), base pixel =
    output pixel - 1, tap offsets, W -> H -> N traversal across rows and images, zero fill of padding pixels, of channel
    slots >= C and of pixels behind the last image."""
    L = _lib()
    b, h, w, c = torch.meshgrid(torch.arange(B), torch.arange(H), torch.arange(W), torch.arange(C), indexing="ij")
    x = ((b + 1) * 1000000 + h * 10000 + w * 100 + c).float().cuda().contiguous()
    out = torch.full((128, cpp), -1.0, device="cuda")
    rc = _probe().tpp_debug_tma_im2col(L.ptr(x), B, H, W, C, cpp, 128, base[0], base[1], base[2], tap[0], tap[1], 0,
                                       L.ptr(out), L.stream_ptr())
    assert rc == 0
    assert torch.equal(out.cpu(), _probe_expected(x, cpp, 128, *base, *tap))


@pytest.mark.parametrize("B,H,W,cin,cout", [(4, 16, 16, 16, 16), (2, 32, 32, 16, 32), (3, 8, 8, 32, 32), (5, 7, 7, 32, 32),
                                            (2, 14, 14, 4, 16), (1, 5, 3, 16, 16), (3, 64, 64, 4, 16),
                                            (5, 32, 32, 32, 16), (40, 16, 16, 32, 32), (300, 32, 32, 16, 16)])
@pytest.mark.parametrize("precision,slots", [(3, 32), (1, 32), (3, 16)])
def test_implicit_conv3x3_matches_conv2d(B, H, W, cin, cout, precision, slots):
    """tpp_gemm_tc in convolution mode (A tiles gathered by TMA im2col from the NHWC TF32 pair) against F.conv2d:
    bias + residual add fused, plain output and ReLU'd TF32 pair output."""
    L = _lib()
    torch.manual_seed(7)
    x = torch.randn(B, H, W, cin, device="cuda")
    w = torch.randn(cout, cin, 3, 3, device="cuda") * 0.1
    bias = torch.randn(cout, device="cuda")
    rows = B * H * W
    skip = torch.randn(rows, cout, device="cuda")
    if slots < cin:
        pytest.skip("16 channel slots per tap need cin <= 16")
    K = 9 * slots
    wf = torch.zeros(cout, 9, slots, device="cuda")
    wf[:, :, :cin] = w.permute(0, 2, 3, 1).reshape(cout, 9, cin)
    wf = wf.view(cout, K)
    xp, wp = _pair(x), _pair(wf)
    out = torch.zeros(rows, cout, device="cuda")
    oh, ol = torch.zeros_like(out), torch.zeros_like(out)
    g = L.TcGemm()
    g.a_hi, g.a_lo = xp[0].data_ptr(), xp[1].data_ptr()
    g.b_hi, g.b_lo, g.ldb = wp[0].data_ptr(), wp[1].data_ptr(), K
    g.M, g.N, g.K, g.precision, g.split_k = rows, cout, K, precision, 1
    g.conv_B, g.conv_H, g.conv_W, g.conv_C = B, H, W, cin
    g.flags = L.EPI_BIAS | L.EPI_ADD | L.EPI_PAIR_RELU
    g.bias, g.addend, g.ld_add = bias.data_ptr(), skip.data_ptr(), cout
    g.out, g.out_hi, g.out_lo, g.ldc = out.data_ptr(), oh.data_ptr(), ol.data_ptr(), cout
    L.call("tpp_gemm_tc", L.C.byref(g), L.stream_ptr())
    ref = F.conv2d(x.double().permute(0, 3, 1, 2), w.double(), bias.double(), padding=1)
    ref = ref.permute(0, 2, 3, 1).reshape(rows, cout) + skip.double()
    _close(out, ref, 2e-5 if precision == 3 else 3e-3)
    assert torch.equal(oh + ol, out.clamp_min(0))


@pytest.mark.parametrize("B,H,W,cin,cout", [(4, 16, 16, 16, 16), (2, 32, 32, 16, 32), (3, 8, 8, 32, 32), (5, 7, 7, 32, 32),
                                            (2, 14, 14, 4, 16), (1, 5, 3, 16, 16)])
@pytest.mark.parametrize("split", [1, 5, 64])
def test_implicit_conv3x3_weight_gradient(B, H, W, cin, cout, split):
    """tpp_gemm_tc in conv_wgrad mode: gw[tap*32 + ci][co] = sum_p X[p + tap][ci] dY[p][co], A blocks gathered by TMA
    im2col (MN-major, 32 pixels per k-block; the last k-block and the last tap tile are partial)."""
    L = _lib()
    torch.manual_seed(8)
    x = torch.randn(B, H, W, cin, device="cuda")
    rows = B * H * W
    dy = torch.randn(rows, cout, device="cuda")
    xp, dyp = _pair(x), _pair(dy)
    gw = torch.zeros(288, cout, device="cuda")
    g = L.TcGemm()
    g.a_hi, g.a_lo = xp[0].data_ptr(), xp[1].data_ptr()
    g.b_hi, g.b_lo, g.ldb = dyp[0].data_ptr(), dyp[1].data_ptr(), cout
    g.M, g.N, g.K, g.precision, g.split_k, g.a_mn, g.b_mn = 288, cout, rows, 3, split, 1, 1
    g.conv_B, g.conv_H, g.conv_W, g.conv_C, g.conv_wgrad = B, H, W, cin, 1
    g.flags, g.out, g.ldc, g.block_n = L.EPI_ACCUM, gw.data_ptr(), cout, 32
    L.call("tpp_gemm_tc", L.C.byref(g), L.stream_ptr())
    xd = x.double().permute(0, 3, 1, 2)
    wd = torch.zeros(cout, cin, 3, 3, dtype=torch.float64, device="cuda", requires_grad=True)
    F.conv2d(xd, wd, padding=1).backward(dy.double().view(B, H, W, cout).permute(0, 3, 1, 2))
    ref = wd.grad.permute(2, 3, 1, 0)                       # [ky][kx][ci][co]
    got = gw.view(3, 3, 32, cout)
    _close(got[:, :, :cin, :], ref, 1e-4)
    assert (got[:, :, cin:, :] == 0).all()


@pytest.mark.parametrize("B,W,cin,cout", [(3, 32, 16, 16), (2, 32, 16, 32), (5, 16, 32, 32), (9, 8, 32, 32),
                                          (160, 32, 16, 16), (700, 8, 32, 32)])
@pytest.mark.parametrize("relu", [False, True])
def test_conv3x3_weight_gradient_fma_kernel(B, W, cin, cout, relu):
    """tpp_conv3x3_wgrad (csrc/conv_cc.cu): exact-fp32 FMA form of the weight gradient for the narrow IMPALA layers,
    gw[(ky*3 + kx)*32 + ci][co] += sum X[y+ky-1][x+kx-1][ci] dY[y][x][co] with X = relu(x), against float64 autograd;
    accumulates into gw (checked by calling it twice); more CTAs than tiles and fewer (persistent loop) both covered."""
    L = _lib()
    torch.manual_seed(B * 100 + W + cin)
    H = W
    x = torch.randn(B, H, W, cin, device="cuda")
    dy = torch.randn(B * H * W, cout, device="cuda")
    gw = torch.zeros(288, cout, device="cuda")
    for _ in range(2):
        L.call("tpp_conv3x3_wgrad", L.ptr(x), 1 if relu else 0, L.ptr(dy), L.ptr(gw), B, H, W, cin, cout, L.stream_ptr())
    xd = (x.clamp_min(0) if relu else x).double().permute(0, 3, 1, 2)
    wd = torch.zeros(cout, cin, 3, 3, dtype=torch.float64, device="cuda", requires_grad=True)
    F.conv2d(xd, wd, padding=1).backward(dy.double().view(B, H, W, cout).permute(0, 3, 1, 2))
    ref = 2 * wd.grad.permute(2, 3, 1, 0)                   # [ky][kx][ci][co], accumulated twice
    got = gw.view(3, 3, 32, cout)
    _close(got[:, :, :cin, :], ref, 2e-6)                   # fp32 FMA chains of <= ~1000 terms + fp32 partial sums
    assert (got[:, :, cin:, :] == 0).all()


def test_conv3x3_weight_gradient_fma_kernel_refuses_other_shapes():
    L = _lib()
    x = torch.randn(2, 14, 14, 16, device="cuda")
    dy = torch.randn(2 * 14 * 14, 16, device="cuda")
    gw = torch.zeros(288, 16, device="cuda")
    rc = L.load().tpp_conv3x3_wgrad(L.ptr(x), 0, L.ptr(dy), L.ptr(gw), 2, 14, 14, 16, 16, L.stream_ptr())
    assert rc == L.ENOTSUP


@pytest.mark.parametrize("B", [1, 3, 40])
def test_first_conv_weight_gradient_fma_kernel(B):
    """tpp_conv3x3_wgrad_first: 3 -> 16 channels on the channel-planar 64 x 64 observation (a row of the minibatch's
    [B, 3*64*64] float matrix, possibly with a padded row stride), gw[(ky*3 + kx)*3 + ci][co], against float64 autograd."""
    L = _lib()
    torch.manual_seed(B)
    H = W = 64
    ld = 3 * H * W + 32
    xbuf = torch.rand(B, ld, device="cuda")
    dy = torch.randn(B * H * W, 16, device="cuda")
    gw = torch.zeros(32, 16, device="cuda")
    L.call("tpp_conv3x3_wgrad_first", L.ptr(xbuf), ld, H * W, W, L.ptr(dy), L.ptr(gw), B, H, W, 16, L.stream_ptr())
    xd = xbuf[:, :3 * H * W].reshape(B, 3, H, W).double()
    wd = torch.zeros(16, 3, 3, 3, dtype=torch.float64, device="cuda", requires_grad=True)
    F.conv2d(xd, wd, padding=1).backward(dy.double().view(B, H, W, 16).permute(0, 3, 1, 2))
    ref = wd.grad.permute(2, 3, 1, 0).reshape(27, 16)        # [(ky, kx, ci)][co]
    _close(gw[:27], ref, 2e-6)
    assert (gw[27:] == 0).all()


@pytest.mark.parametrize("B", [1, 5, 64])
def test_first_conv_forward_fma_kernel(B):
    """tpp_conv3x3_fwd_first against torch conv2d (float64) on the planar observation, NHWC output, torch weight layout."""
    L = _lib()
    torch.manual_seed(B + 7)
    H = W = 64
    ld = 3 * H * W + 16
    xbuf = torch.rand(B, ld, device="cuda")
    w = torch.randn(16, 3, 3, 3, device="cuda") * 0.3
    bias = torch.randn(16, device="cuda")
    out = torch.zeros(B, H, W, 16, device="cuda")
    L.call("tpp_conv3x3_fwd_first", L.ptr(xbuf), ld, H * W, W, L.ptr(w), L.ptr(bias), L.ptr(out), B, H, W, 16,
           L.stream_ptr())
    ref = F.conv2d(xbuf[:, :3 * H * W].reshape(B, 3, H, W).double(), w.double(), bias.double(), padding=1)
    _close(out, ref.permute(0, 2, 3, 1), 1e-6)


@pytest.mark.parametrize("B,cin,cout", [(1, 16, 16), (3, 16, 32), (2, 32, 16), (70, 16, 16)])
@pytest.mark.parametrize("mode", ["fwd_pair_relu", "fwd_add", "dgrad_mask_add_colsum", "plain_only"])
def test_conv3x3_fma_forward_and_data_gradient_forms(B, cin, cout, mode):
    """tpp_conv3x3_fma (csrc/conv_cc.cu) against torch conv2d in float64, with the epilogue variants the engine uses:
    forward (bias, ReLU'd input, residual add, plain + TF32 pair of relu(y)) and data gradient (ReLU mask, skip-gradient
    add, bias-gradient column sums).  GEMM-layout weights wg[co][tap*slots + ci], slots = 16 or 32."""
    L = _lib()
    torch.manual_seed(B * 31 + cin + len(mode))
    H = W = 32
    rows = B * H * W
    slots = 16 if cin <= 16 else 32
    x = torch.randn(B, H, W, cin, device="cuda")
    w = torch.randn(cout, cin, 3, 3, device="cuda") * 0.2
    wg = torch.zeros(cout, 9 * slots, device="cuda")
    wg.view(cout, 3, 3, slots)[..., :cin] = w.permute(0, 2, 3, 1)
    bias = torch.randn(cout, device="cuda") if mode.startswith("fwd") else None
    relu_in = mode.startswith("fwd")
    mask = torch.randn(rows, cout, device="cuda") if "mask" in mode else None
    addend = torch.randn(rows, cout, device="cuda") if "add" in mode else None
    pair_relu = mode == "fwd_pair_relu"
    want_pair = mode != "plain_only"
    colsum = torch.full((cout,), 0.5, device="cuda") if "colsum" in mode else None
    out = torch.zeros(rows, cout, device="cuda")
    oh, ol = (torch.zeros_like(out), torch.zeros_like(out)) if want_pair else (None, None)
    L.call("tpp_conv3x3_fma", L.ptr(x), 1 if relu_in else 0, L.ptr(wg), slots, L.ptr(bias), L.ptr(mask), L.ptr(addend),
           1 if pair_relu else 0, L.ptr(out), L.ptr(oh), L.ptr(ol), L.ptr(colsum), B, H, W, cin, cout, L.stream_ptr())
    xin = (x.clamp_min(0) if relu_in else x).double().permute(0, 3, 1, 2)
    ref = F.conv2d(xin, w.double(), bias.double() if bias is not None else None, padding=1)
    ref = ref.permute(0, 2, 3, 1).reshape(rows, cout)
    if mask is not None:
        ref = ref * (mask > 0)
    if addend is not None:
        ref = ref + addend.double()
    _close(out, ref, 2e-6)
    if want_pair:
        y = out.clamp_min(0) if pair_relu else out
        assert torch.equal(oh + ol, y)                                     # the pair is an exact split of the fp32 value
        assert torch.equal(oh, ((y.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32))
    if colsum is not None:
        _close(colsum - 0.5, ref.sum(0), 1e-5)


def test_fma_conv_kernels_at_the_full_minibatch_size():
    """The FMA-pipe kernels at the Procgen minibatch's size (2048 frames: 2.1 M pixels at 32 x 32, 8.4 M at 64 x 64 -- the
    persistent tile loop, 64-bit offsets, every CTA's partial sums), through properties that do not need a float64
    convolution of that size: (1) the centre tap of the weight gradient is the plain contraction X^T dY; (2) corner taps
    equal the contraction of shifted views; (3) linearity in dY; (4) agreement with the tensor-core form of the same
    contraction; (5) first-layer forward against torch's fp32 convolution."""
    L = _lib()
    torch.manual_seed(11)
    B, H, W, Cc = 2048, 32, 32, 16
    rows = B * H * W
    x = torch.randn(B, H, W, Cc, device="cuda")
    dy1, dy2 = torch.randn(rows, Cc, device="cuda"), torch.randn(rows, Cc, device="cuda")

    def wgrad(dy, relu=1):
        gw = torch.zeros(288, Cc, device="cuda")
        L.call("tpp_conv3x3_wgrad", L.ptr(x), relu, L.ptr(dy), L.ptr(gw), B, H, W, Cc, Cc, L.stream_ptr())
        return gw.view(3, 3, 32, Cc)[:, :, :Cc, :]

    g1, g2 = wgrad(dy1), wgrad(dy2)
    xr = x.clamp_min(0)
    centre = torch.einsum("pi,pj->ij", xr.view(rows, Cc).double(), dy1.double())
    _close(g1[1, 1], centre, 2e-6)
    d4 = dy1.view(B, H, W, Cc)
    corner = torch.einsum("bhwi,bhwj->ij", xr[:, :-1, :-1].double(), d4[:, 1:, 1:].double())     # tap (0, 0): X[y-1][x-1]
    _close(g1[0, 0], corner, 2e-6)
    corner = torch.einsum("bhwi,bhwj->ij", xr[:, 1:, :-1].double(), d4[:, :-1, 1:].double())      # tap (2, 0): X[y+1][x-1]
    _close(g1[2, 0], corner, 2e-6)
    _close(wgrad((dy1 + 2 * dy2).contiguous()), g1.double() + 2 * g2.double(), 2e-6)
    # the tensor-core form (TMA im2col, 3xTF32) on the same operands
    xp, dyp = _pair(xr.contiguous()), _pair(dy1)
    gtc = torch.zeros(288, Cc, device="cuda")
    g = L.TcGemm()
    g.a_hi, g.a_lo = xp[0].data_ptr(), xp[1].data_ptr()
    g.b_hi, g.b_lo, g.ldb = dyp[0].data_ptr(), dyp[1].data_ptr(), Cc
    g.M, g.N, g.K, g.precision, g.split_k, g.a_mn, g.b_mn = 288, Cc, rows, 3, rows // 1024, 1, 1
    g.conv_B, g.conv_H, g.conv_W, g.conv_C, g.conv_wgrad = B, H, W, Cc, 1
    g.flags, g.out, g.ldc, g.block_n = L.EPI_ACCUM, gtc.data_ptr(), Cc, 32
    L.call("tpp_gemm_tc", L.C.byref(g), L.stream_ptr())
    _close(gtc.view(3, 3, 32, Cc)[:, :, :Cc, :], g1.double(), 1e-4)
    # first convolution at 2048 frames of 64 x 64
    xb = torch.rand(B, 3 * 64 * 64, device="cuda")
    w = torch.randn(16, 3, 3, 3, device="cuda") * 0.3
    bias = torch.randn(16, device="cuda")
    out = torch.zeros(B, 64, 64, 16, device="cuda")
    L.call("tpp_conv3x3_fwd_first", L.ptr(xb), 3 * 64 * 64, 64 * 64, 64, L.ptr(w), L.ptr(bias), L.ptr(out), B, 64, 64, 16,
           L.stream_ptr())
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        ref = F.conv2d(xb.view(B, 3, 64, 64), w, bias, padding=1).permute(0, 2, 3, 1)
    finally:
        torch.backends.cudnn.allow_tf32 = prev
    _close(out, ref, 3e-6)
    dyf = torch.randn(B * 64 * 64, 16, device="cuda")
    gwf = torch.zeros(32, 16, device="cuda")
    L.call("tpp_conv3x3_wgrad_first", L.ptr(xb), 3 * 64 * 64, 64 * 64, 64, L.ptr(dyf), L.ptr(gwf), B, 64, 64, 16,
           L.stream_ptr())
    centre = torch.einsum("bcp,bpj->cj", xb.view(B, 3, 4096).double(), dyf.view(B, 4096, 16).double())
    _close(gwf[:27].view(3, 3, 3, 16)[1, 1], centre, 2e-6)
