"""GPU parity tests of the RDNet companion kernels (depthwise 7x7 + LayerNorm2d, LayerNorm2d with space-to-depth, the
EffectiveSE gate; pssr/models/_rdnet.py:57-62, :172-174, :181-183, :200-202), each through the plan API / C ABI against a
plain PyTorch fp32 statement of the same op on the same 16-bit operands.  Tolerance: one rounding of the fp16 output
(2^-11 relative) plus fp32 summation-order noise; (hi, lo) pair outputs are checked three orders of magnitude tighter."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _P():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    from pssr2_b200 import plan as P
    return P


def _act(shape, seed, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*shape, device="cuda", generator=g) * scale)


def _pair(x):
    hi = x.half()
    return hi, (x - hi.float()).half()


def _nchw(v):
    return v.float().permute(0, 3, 1, 2).contiguous()


def _ln2d(x, w, b, eps):       # LayerNorm over the channel axis of an NCHW tensor (timm LayerNorm2d)
    return F.layer_norm(x.permute(0, 2, 3, 1), (x.shape[1],), w, b, eps).permute(0, 3, 1, 2)


def _close(got, ref, rel, what):
    err = (got - ref).abs()
    tol = rel * ref.abs() + rel * 4 * max(1.0, float(ref.abs().max())) * 2e-3 + 1e-4 * (rel / 2.0 ** -11)
    bad = err > tol
    assert not bool(bad.any()), f"{what}: {int(bad.sum())} / {bad.numel()} out of tolerance, max err {float(err.max()):.3g} (ref max {float(ref.abs().max()):.3g})"


@pytest.mark.parametrize("B,H,W,C,choff", [(3, 16, 16, 128, 0), (2, 20, 24, 72, 8), (2, 8, 8, 704, 0), (1, 64, 64, 64, 0), (5, 13, 7, 200, 0),
                                              (10, 64, 128, 136, 0)])      # the last one: the slab-walking (cp.async pipelined) kernel
def test_dwconv7_layernorm(B, H, W, C, choff):
    P = _P()
    buf = _act((B, H, W, C + choff + 8), 1).half().contiguous()          # the source is a channel slice of a wider buffer
    src = P.View(buf, choff, C)
    dw = _act((C, 1, 7, 7), 2, 0.15)
    db, lw, lb = _act((C,), 3, 0.1), 1.0 + _act((C,), 4, 0.1), _act((C,), 5, 0.1)
    out = torch.zeros(B, H, W, C, dtype=torch.float16, device="cuda")
    plan = P.Plan("fp16")
    plan.dwconv_ln(src, dw.view(C, 49).t().contiguous(), db, lw, lb, 1e-6, P.View(out))
    plan.finalize().run()
    torch.cuda.synchronize()
    x = _nchw(buf[..., choff:choff + C])
    pre = F.conv2d(x, dw, db, padding=3, groups=C)
    # the kernel stores the pre-LayerNorm value as fp16 before normalising it (two launches): same rounding here
    ref = _ln2d(pre.half().float(), lw, lb, 1e-6)
    _close(_nchw(out), ref, 2.0 ** -11 * 4, f"dwconv7+ln {B}x{H}x{W}x{C}")


@pytest.mark.parametrize("B,H,W,C", [(2, 16, 16, 128), (1, 24, 40, 192)])
def test_dwconv7_layernorm_pairs(B, H, W, C):
    """Compensated precision: the input and the output travel as (hi, lo) fp16 pairs; the result carries ~22 bits."""
    P = _P()
    x32 = _act((B, H, W, C), 11)
    hi, lo = _pair(x32)
    dw = _act((C, 1, 7, 7), 12, 0.15)
    db, lw, lb = _act((C,), 13, 0.1), 1.0 + _act((C,), 14, 0.1), _act((C,), 15, 0.1)
    out, out_lo = torch.zeros_like(hi), torch.zeros_like(hi)
    plan = P.Plan("fp16")
    plan.dwconv_ln(P.View(hi.contiguous()), dw.view(C, 49).t().contiguous(), db, lw, lb, 1e-6, P.View(out), src_lo=P.View(lo.contiguous()),
                   out_lo=P.View(out_lo))
    plan.finalize().run()
    torch.cuda.synchronize()
    x = _nchw(hi) + _nchw(lo)
    ref = _ln2d(F.conv2d(x.double(), dw.double(), db.double(), padding=3, groups=C), lw.double(), lb.double(), 1e-6).float()
    got = _nchw(out) + _nchw(out_lo)
    err = float((got - ref).abs().max())
    single = float((_nchw(out) - ref).abs().max())
    assert err <= 2e-5 * max(1.0, float(ref.abs().max())), (err, single)
    assert single > 4 * err           # the lo half carries real information


@pytest.mark.parametrize("B,H,W,C,s2d", [(2, 16, 16, 320, 2), (3, 8, 12, 64, 1), (1, 8, 8, 1040, 2), (2, 10, 6, 520, 1)])
def test_layernorm_space_to_depth(B, H, W, C, s2d):
    P = _P()
    x = _act((B, H, W, C), 21, 2.0).half().contiguous()
    lw, lb = 1.0 + _act((C,), 22, 0.1), _act((C,), 23, 0.1)
    out = torch.zeros(B, H // s2d, W // s2d, C * s2d * s2d, dtype=torch.float16, device="cuda")
    plan = P.Plan("fp16")
    plan.layernorm(P.View(x), lw, lb, 1e-6, P.View(out), s2d=s2d)
    plan.finalize().run()
    torch.cuda.synchronize()
    ref = _ln2d(_nchw(x), lw, lb, 1e-6)                                   # [B, C, H, W]
    if s2d == 2:     # channel block (dy * 2 + dx) of output pixel (y / 2, x / 2) holds input pixel (y, x): a 2x2 stride-2 conv becomes 1x1
        ref = ref.view(B, C, H // 2, 2, W // 2, 2).permute(0, 3, 5, 1, 2, 4).reshape(B, 4 * C, H // 2, W // 2)
    _close(_nchw(out), ref, 2.0 ** -11 * 2, f"layernorm s2d={s2d} C={C}")


@pytest.mark.parametrize("B,H,W,C,with_gamma", [(5, 16, 16, 128, True), (3, 8, 8, 224, True), (2, 16, 16, 64, False), (2, 32, 32, 64, True),
                                                (4, 4, 4, 512, True)])
def test_effective_se(B, H, W, C, with_gamma):
    """Single-launch kernel on small maps (the first, second, third and fifth case), two-pass fallback on the 32x32 map."""
    P = _P()
    x = _act((B, H, W, C), 31).half().contiguous()
    fw, fb = _act((C, C), 32, 1.0 / C ** 0.5), _act((C,), 33, 0.5)
    gamma = (0.5 + _act((C,), 34, 0.2)) if with_gamma else None
    out = torch.zeros(B, H, W, C + 16, dtype=torch.float16, device="cuda")
    gate_ws = torch.zeros(B * C, dtype=torch.float32, device="cuda")
    plan = P.Plan("fp16")
    plan.ese(P.View(x), fw.contiguous(), fb, gamma, gate_ws, P.View(out, 8, C))
    plan.finalize().run()
    torch.cuda.synchronize()
    xf = _nchw(x)
    se = F.conv2d(xf.mean((2, 3), keepdim=True), fw.view(C, C, 1, 1), fb)
    ref = xf * (F.relu6(se + 3.0) / 6.0)
    if gamma is not None:
        ref = ref * gamma.view(1, -1, 1, 1)
    _close(_nchw(out[..., 8:8 + C]), ref, 2.0 ** -11 * 2, f"eSE {B}x{H}x{W}x{C}")
    assert float(out[..., :8].abs().max()) == 0.0 and float(out[..., 8 + C:].abs().max()) == 0.0     # the slice's neighbours are untouched


def test_table_fetch_matches_memcpy():
    """pssr_table_fetch: the one-CTA kernel reads a pinned host buffer in place and lands the same bytes a copy would."""
    from pssr2_b200 import _lib
    host = torch.randint(0, 255, (4096 + 52,), dtype=torch.uint8).pin_memory()
    dst = torch.zeros(host.numel(), dtype=torch.uint8, device="cuda")
    _lib.check(_lib.lib().pssr_table_fetch(dst.data_ptr(), host.data_ptr(), host.numel(), _lib.current_stream_ptr(dst.device)))
    torch.cuda.synchronize()
    assert torch.equal(dst.cpu(), host)
    assert _lib.lib().pssr_table_fetch(dst.data_ptr(), host.data_ptr(), 3, _lib.current_stream_ptr(dst.device)) != 0      # bytes % 4
    # pageable source: falls back to cudaMemcpyAsync, same bytes
    pageable = torch.randint(0, 255, (256,), dtype=torch.uint8)
    _lib.check(_lib.lib().pssr_table_fetch(dst.data_ptr(), pageable.data_ptr(), 256, _lib.current_stream_ptr(dst.device)))
    torch.cuda.synchronize()
    assert torch.equal(dst[:256].cpu(), pageable)
