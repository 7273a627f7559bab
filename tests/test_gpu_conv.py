"""GPU parity tests for the tcgen05 implicit-GEMM convolution and its companions, through the C ABI.

Floating-point kernel => the comparison is against a plain PyTorch fp32 reference of the same op on
the same 16-bit-rounded operands (TF32 disabled); tolerance = one rounding of the 16-bit output
format (bf16: 2^-8 relative, fp16: 2^-11 relative) plus fp32 accumulation-order noise.
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

REL = {"bf16": 2.0 ** -8, "fp16": 2.0 ** -11}


def _setup():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    from pssr2_b200 import plan as P
    return P


def _rand_act(B, H, W, C, dt, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(B, H, W, C, device="cuda", generator=g)).to(dt).contiguous()


def _nchw(v):  # NHWC 16-bit -> NCHW fp32
    return v.float().permute(0, 3, 1, 2).contiguous()


def _check(out16, ref, prec, what):
    got = out16.float()
    err = (got - ref).abs()
    tol = REL[prec] * ref.abs() + 2e-3 * max(1.0, float(ref.abs().max())) * REL[prec] * 4 + 1e-4
    bad = err > tol
    assert not bool(bad.any()), (f"{what}: {int(bad.sum())} / {bad.numel()} elements out of tolerance, "
                                 f"max err {float(err.max()):.5g}, max ref {float(ref.abs().max()):.5g}")


CASES = [
    # B, H, W, Cin, Cout, shuffle, prec
    (2, 32, 32, 64, 64, 1, "bf16"),
    (2, 16, 16, 128, 256, 1, "bf16"),
    (4, 8, 8, 256, 512, 1, "bf16"),
    (16, 4, 4, 64, 64, 1, "bf16"),
    (1, 20, 20, 64, 64, 1, "bf16"),      # ragged: tiles overhang the image
    (3, 16, 16, 96, 64, 1, "bf16"),      # ragged K: second 64-channel block is half out of bounds
    (2, 16, 16, 64, 128, 2, "bf16"),     # pixel_shuffle(2) epilogue
    (1, 16, 16, 64, 1024, 4, "bf16"),    # pixel_shuffle(4), 4 N tiles
    (8, 64, 64, 64, 64, 1, "bf16"),      # 256 tiles > 148 SMs: persistent loop + TMEM double buffering
    (2, 32, 32, 64, 64, 1, "fp16"),
    (2, 16, 16, 128, 256, 2, "fp16"),
]


@pytest.mark.parametrize("B,H,W,Cin,Cout,shuffle,prec", CASES)
def test_conv3x3(B, H, W, Cin, Cout, shuffle, prec):
    P = _setup()
    plan = P.Plan(prec)
    dt = plan.tdtype
    x = _rand_act(B, H, W, Cin, dt, 1)
    g = torch.Generator(device="cuda").manual_seed(2)
    w = torch.randn(Cout, Cin, 3, 3, device="cuda", generator=g) / (3.0 * Cin ** 0.5)
    b = torch.randn(Cout, device="cuda", generator=g)
    wp = P.pack_weight([w], plan.dtype, shuffle)
    bp = P.permute_n(b, shuffle).contiguous()
    cps = Cout // (shuffle * shuffle)
    out = torch.full((B, H * shuffle, W * shuffle, cps + 8), 7.0, dtype=dt, device="cuda")  # written at channel offset 8
    plan.conv([P.View(x)], [(0, 9, P.ceil_div(Cin, 64))], wp, bp, P.View(out, 8, cps), Ho=H, Wo=W, B=B, shuffle=shuffle,
              act=P.ACT_RELU)
    plan.finalize()
    plan.run()
    torch.cuda.synchronize()
    ref = F.relu(F.conv2d(_nchw(x), w.to(dt).float(), b, padding=1))
    if shuffle > 1:
        ref = F.pixel_shuffle(ref, shuffle)
    _check(_nchw(out[..., 8:]), ref, prec, f"conv3x3 {B}x{H}x{W} {Cin}->{Cout} r={shuffle} {prec}")
    assert bool((out[..., :8].float() == 7.0).all()), "conv wrote outside its channel slice"


def test_conv_residual_two_sources():
    """relu(conv3x3(h) + conv1x1(x) + bias): the fused ResBlock tail (pssr/models/_blocks.py:39-41)."""
    P = _setup()
    plan = P.Plan("bf16")
    dt = plan.tdtype
    B, H, W, C, Cx = 2, 16, 16, 128, 192
    h = _rand_act(B, H, W, C, dt, 3)
    xbuf = _rand_act(B, H, W, Cx + 64, dt, 4)          # x is a channel slice [64, 64+Cx) of a wider buffer
    g = torch.Generator(device="cuda").manual_seed(5)
    w3 = torch.randn(C, C, 3, 3, device="cuda", generator=g) / (3.0 * C ** 0.5)
    w1 = torch.randn(C, Cx, 1, 1, device="cuda", generator=g) / (Cx ** 0.5)
    b = torch.randn(C, device="cuda", generator=g)
    wp = P.pack_weight([w3, w1], plan.dtype)
    out = torch.zeros(B, H, W, C, dtype=dt, device="cuda")
    plan.conv([P.View(h), P.View(xbuf, 64, Cx)], [(0, 9, 2), (1, 1, 3)], wp, b, P.View(out), Ho=H, Wo=W, B=B, act=P.ACT_RELU)
    plan.finalize()
    plan.run()
    torch.cuda.synchronize()
    ref = F.relu(F.conv2d(_nchw(h), w3.to(dt).float(), None, padding=1) + F.conv2d(_nchw(xbuf[..., 64:]), w1.to(dt).float(), b))
    _check(_nchw(out), ref, "bf16", "conv+respass")


def test_prep_pool_tail():
    P = _setup()
    from pssr2_b200 import models as M
    plan = P.Plan("bf16")
    dt = plan.tdtype
    B, C, H, W = 2, 1, 32, 32
    g = torch.Generator(device="cuda").manual_seed(6)
    x = torch.randint(0, 256, (B, C, H, W), device="cuda", generator=g).float()
    sc = torch.tensor([1.3], device="cuda")
    sh = torch.tensor([-0.2], device="cuda")
    im2col = torch.zeros(B, H, W, 64, dtype=dt, device="cuda")
    plan.prep(x, sc, sh, im2col)
    src = _rand_act(B, H, W, 72, dt, 7)
    pooled = torch.zeros(B, H // 2, W // 2, 64, dtype=dt, device="cuda")
    plan.maxpool(P.View(src, 8, 64), P.View(pooled))
    tin = _rand_act(B, H, W, 64, dt, 8)
    wt = torch.randn(1, 64, 3, 3, device="cuda", generator=g) / 24.0
    bt = torch.randn(1, device="cuda", generator=g)
    out = torch.zeros(B, 1, H, W, device="cuda")
    out8 = torch.zeros(B, 1, H, W, dtype=torch.uint8, device="cuda")
    plan.tail(P.View(tin), wt.permute(0, 2, 3, 1).contiguous(), bt, 128.0, 128.0, out, out8)
    plan.finalize()
    plan.run()
    torch.cuda.synchronize()
    xn = (x / 128 - 1) * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1)
    cols = F.unfold(xn, 3, padding=1).view(B, C * 9, H, W).to(dt)
    assert torch.equal(_nchw(im2col)[:, :9 * C], cols.float()), "im2col mismatch"
    assert bool((im2col[..., 9 * C:].float() == 0).all())
    assert torch.equal(_nchw(pooled), F.max_pool2d(_nchw(src[..., 8:]), 2)), "maxpool mismatch"
    ref = F.conv2d(_nchw(tin), wt, bt, padding=1) * 128 + 128
    assert float((out - ref).abs().max()) < 2e-3, f"tail max err {float((out - ref).abs().max())}"
    ref8 = out.clamp(0, 255).to(torch.uint8)
    assert torch.equal(out8, ref8), "tail uint8 truncation mismatch"
