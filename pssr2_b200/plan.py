"""Host-side builder for a network plan (family 2 of ``include/pssr_b200.h``).

The modules in ``pssr2_b200/models.py`` describe their forward pass (pssr/models/resunet.py:65-96,
rdresunet.py:104-130) as a list of fused ops over NHWC 16-bit activation buffers; this module packs
weights into the K-major layout the tcgen05 kernel reads, fills the C structs and owns the
``pssr_plan_t`` handle.  torch is used for device memory only.
"""
import ctypes
import math

import torch

from . import _lib
from ._lib import (ACT_GELU, ACT_NONE, ACT_RELU, DT_BF16, DT_FP16, OP_CAST8, OP_CONV, OP_DWCONV_LN, OP_ESE, OP_LAYERNORM, OP_MAXPOOL,
                   OP_PREP, OP_RESAMPLE, OP_STEM, OP_TAIL, OP_TAILSUM, OP_WINATTN, SEG_E5M2, SEG_F16, Cast8Desc, ConvDesc, DwLnDesc, EseDesc, KSeg, LnDesc, Op,
                   PoolDesc, PrepDesc, ResampleDesc, Src, StemDesc, TailDesc, TailSumDesc, WinAttnDesc)

TORCH_DT = {DT_BF16: torch.bfloat16, DT_FP16: torch.float16}
DT_NAMES = {"bf16": DT_BF16, "fp16": DT_FP16, "fp16c": DT_FP16}
TAIL_COMP = 1   # include/pssr_b200.h PSSR_TAIL_COMP
__all_seg__ = (SEG_F16, SEG_E5M2)


DRY_RUN = False


def ceil_div(a, b):
    return -(-a // b)


class View:
    """A channel slice [choff, choff+channels) of an NHWC 16-bit buffer [B, H, W, cstride]."""

    def __init__(self, buf: torch.Tensor, choff: int = 0, channels: int = None):
        assert buf.dim() == 4 and buf.is_contiguous()
        self.buf = buf
        self.choff = choff
        self.channels = buf.shape[3] - choff if channels is None else channels
        assert self.choff % 8 == 0 and self.choff + self.channels <= buf.shape[3]

    @property
    def B(self):
        return self.buf.shape[0]

    @property
    def H(self):
        return self.buf.shape[1]

    @property
    def W(self):
        return self.buf.shape[2]

    @property
    def cstride(self):
        return self.buf.shape[3]

    def ptr(self):
        return self.buf.data_ptr() + self.choff * self.buf.element_size()


def split_lo(w, dtype):
    """What the 16-bit rounding of ``w`` drops: w - rn16(w) (fp32; packed as its own K blocks by the compensated layers)."""
    return w - w.to(TORCH_DT[dtype]).float()


def pack_weight(parts, dtype, shuffle=1, n_pad=None):
    """parts: list of fp32 tensors [Cout, Cin, kh, kw] in K-schedule order (one per K segment).
    Returns ([n_pad, Ktot] 16-bit, K-major, K ordered tap-major then 64-padded channel) where the N
    rows are permuted so that F.pixel_shuffle(., shuffle) becomes a contiguous store:
    new n = (i*r + j)*C' + c'  <-  old c = c'*r*r + i*r + j   (torch pixel_shuffle channel order)."""
    cols = []
    cout = parts[0].shape[0]
    for w in parts:
        co, ci, kh, kw = w.shape
        assert co == cout
        cb = ceil_div(ci, 64)
        wp = torch.zeros(co, kh * kw, cb * 64, dtype=torch.float32, device=w.device)
        wp[:, :, :ci] = w.permute(0, 2, 3, 1).reshape(co, kh * kw, ci)
        cols.append(wp.reshape(co, -1))
    W = torch.cat(cols, 1)
    W = permute_n(W, shuffle)
    n_pad = cout if n_pad is None else n_pad
    if n_pad > cout:
        W = torch.cat([W, torch.zeros(n_pad - cout, W.shape[1], device=W.device)], 0)
    return W.to(TORCH_DT[dtype]).contiguous()


def pack_weight8(parts, scale, shuffle=1):
    """e5m2 twin of ``pack_weight`` for the PSSR_SEG_E5M2 segments: [n, K8tot] uint8 holding e5m2(w * scale), same K / N order."""
    cols = []
    for w in parts:
        co, ci, kh, kw = w.shape
        cb = ceil_div(ci, 64)
        wp = torch.zeros(co, kh * kw, cb * 64, dtype=torch.float32, device=w.device)
        wp[:, :, :ci] = w.permute(0, 2, 3, 1).reshape(co, kh * kw, ci)
        cols.append(wp.reshape(co, -1))
    W = permute_n(torch.cat(cols, 1) * scale, shuffle)
    return W.clamp(-57344.0, 57344.0).to(torch.float8_e5m2).view(torch.uint8).contiguous()


def permute_n(t, shuffle):
    """Reorders dim 0 (output channels) for the pixel-shuffle epilogue."""
    if shuffle == 1:
        return t
    r2 = shuffle * shuffle
    c = t.shape[0]
    assert c % r2 == 0
    rest = t.shape[1:]
    return t.reshape(c // r2, r2, *rest).transpose(0, 1).reshape(c, *rest)


class Plan:
    def __init__(self, dtype="bf16"):
        # "fp16c": fp16 operands with hi + lo compensation on the layers the output is sensitive to (models.py)
        self.comp = dtype == "fp16c"
        self.dtype = DT_NAMES[dtype] if isinstance(dtype, str) else dtype
        self.tdtype = TORCH_DT[self.dtype]
        self.ops = []
        self.keep = []      # tensors that must outlive the plan
        self.handle = None
        self.flops = 0      # algorithmic conv FLOPs (2*MAC, unpadded)
        self.records = []   # python-level mirror of the op list (used by the CPU plan interpreter in tests/)

    # ---- op builders -----------------------------------------------------------------
    def conv(self, srcs, segs, weight, bias, out: View, *, Ho, Wo, B, n=None, n_valid=None, shuffle=1, act=ACT_NONE,
             out_scale=None, out_f32=None, tail_weight=None, tail_z=None, tail_layout=0, out_lo: View = None, tail_flags=0, weight8=None,
             resid: View = None, resid_scale=1.0):
        """srcs: list[View]; segs: list[(src_index, taps, cblocks[, fmt[, dilation]])]; weight [n, Ktot] 16-bit; bias [n] fp32;
        weight8 [n, K8tot] uint8 (e5m2) for the segments with fmt == SEG_E5M2, whose sources are uint8 (e5m2) NHWC views."""
        d = ConvDesc()
        d.n_srcs = len(srcs)
        for i, s in enumerate(srcs):
            d.srcs[i] = Src(s.ptr(), s.channels, s.cstride, s.H, s.W, s.B, 0)
        assert len(srcs) <= 4 and len(segs) <= 6
        d.n_segs = len(segs)
        segs = [(tuple(sg) + (SEG_F16, 1)[len(sg) - 3:]) for sg in segs]
        ktot = k8tot = 0
        for i, (si, taps, cb, fmt, dil) in enumerate(segs):
            d.segs[i] = KSeg(si, taps, cb, fmt, dil)
            assert dil >= 1 and (dil == 1 or taps == 9)
            assert (srcs[si].buf.dtype == torch.uint8) == (fmt == SEG_E5M2)
            if fmt == SEG_E5M2:
                k8tot += taps * cb * 64
            else:
                ktot += taps * cb * 64
        n = weight.shape[0] if n is None else n
        assert weight.shape == (n, ktot) and weight.dtype == self.tdtype and weight.is_contiguous()
        assert (weight8 is None) == (k8tot == 0)
        if weight8 is not None:
            assert weight8.shape == (n, k8tot) and weight8.dtype == torch.uint8 and weight8.is_contiguous()
            d.weights8 = weight8.data_ptr()
        assert bias.shape == (n,) and bias.dtype == torch.float32
        d.weights = weight.data_ptr()
        d.bias = bias.data_ptr()
        d.n = n
        d.n_valid = n if n_valid is None else n_valid
        d.Ho, d.Wo, d.B = Ho, Wo, B
        d.out = out.buf.data_ptr() if out is not None else None
        d.out_cstride = out.cstride if out is not None else (out_f32.shape[3] if out_f32 is not None else 16)
        d.out_choff = out.choff if out is not None else 0
        d.shuffle = shuffle
        d.act = act
        d.out_scale = out_scale.data_ptr() if out_scale is not None else None
        d.out_f32 = out_f32.data_ptr() if out_f32 is not None else None
        d.tail_weight = tail_weight.data_ptr() if tail_weight is not None else None
        d.tail_z = tail_z.data_ptr() if tail_z is not None else None
        d.tail_layout = tail_layout
        d.tail_flags = tail_flags
        if out_lo is not None:
            assert out is not None and (out_lo.B, out_lo.H, out_lo.W) == (out.B, out.H, out.W)
            d.out_lo = out_lo.buf.data_ptr()
            d.out_lo_cstride = out_lo.cstride
            d.out_lo_choff = out_lo.choff
        if resid is not None:
            # (with a pixel shuffle the residual is read in the GEMM's own column order: pack its producer with the same shuffle)
            assert (resid.B, resid.H, resid.W) == (B, Ho, Wo) and resid.buf.dtype == self.tdtype and resid.channels >= d.n_valid
            d.resid = resid.ptr()
            d.resid_cstride = resid.cstride
            d.resid_choff = 0
            d.resid_scale = float(resid_scale)
        op = Op()
        op.kind = OP_CONV
        op.u.conv = d
        self.ops.append(op)
        self.keep += [weight, bias, out_scale, out_f32, tail_weight, tail_z] + [s.buf for s in srcs] + ([out.buf] if out is not None else [])
        self.keep += [out_lo.buf] if out_lo is not None else []
        self.keep += [weight8, resid.buf if resid is not None else None]
        self.records.append(("conv", dict(srcs=list(srcs), segs=list(segs), weight=weight, bias=bias, out=out, Ho=Ho, Wo=Wo, B=B,
                                          n=n, n_valid=d.n_valid, shuffle=shuffle, act=act, out_scale=out_scale, out_f32=out_f32,
                                          issued_flops=2 * B * Ho * Wo * n * ktot, tail_weight=tail_weight, tail_z=tail_z,
                                          tail_layout=tail_layout, out_lo=out_lo, tail_flags=tail_flags, weight8=weight8, resid=resid,
                                          resid_scale=float(resid_scale))))

    def prep(self, x, scale, shift, im2col, xnorm=None, centre_only=False, im2col_lo=None):
        """centre_only: im2col is the normalised input itself, NHWC [B, H, W, cols] (inputs with more than 7 channels).
        im2col_lo (compensated precision): same shape, receives what the 16-bit rounding of every im2col value dropped."""
        B, C, H, W = x.shape
        assert im2col_lo is None or (im2col_lo.shape == im2col.shape and im2col_lo.dtype == im2col.dtype)
        d = PrepDesc(x.data_ptr(), 1 if x.dtype == torch.uint8 else 0, B, C, H, W, scale.data_ptr(), shift.data_ptr(),
                     im2col.data_ptr(), im2col_lo.data_ptr() if im2col_lo is not None else None,
                     xnorm.data_ptr() if xnorm is not None else None, im2col.shape[3], 1 if centre_only else 0)
        if centre_only:
            assert im2col.shape[3] % 8 == 0 and C <= im2col.shape[3] <= 64
        else:
            assert im2col.shape[3] in (16, 64) and C * 9 <= im2col.shape[3]
        op = Op()
        op.kind = OP_PREP
        op.u.prep = d
        self.ops.append(op)
        self.keep += [x, scale, shift, im2col, xnorm, im2col_lo]
        self.records.append(("prep", dict(x=x, scale=scale, shift=shift, im2col=im2col, xnorm=xnorm, centre_only=centre_only,
                                          im2col_lo=im2col_lo)))

    def cast8(self, src: View, dst: View, scale):
        """dst (uint8 NHWC, e5m2 bit patterns) = e5m2(src * scale): the operand of SEG_E5M2 segments."""
        assert dst.buf.dtype == torch.uint8 and src.channels == dst.channels and src.channels % 16 == 0
        d = Cast8Desc(src.buf.data_ptr(), src.cstride, src.choff, src.channels, src.B, src.H, src.W, float(scale), dst.buf.data_ptr(),
                      dst.cstride, dst.choff)
        op = Op()
        op.kind = OP_CAST8
        op.u.cast8 = d
        self.ops.append(op)
        self.keep += [src.buf, dst.buf]
        self.records.append(("cast8", dict(src=src, dst=dst, scale=float(scale))))

    def _resample(self, src: View, dst: View, mode, k=1, relu=False, scale=None, shift=None):
        Ho, Wo = (src.H, src.W) if mode == 0 else (dst.H, dst.W)
        assert src.channels == dst.channels and src.channels % 8 == 0 and src.B == dst.B and (dst.H, dst.W) == (Ho, Wo)
        d = ResampleDesc(src.buf.data_ptr(), src.cstride, src.choff, src.channels, src.B, src.H, src.W, mode, k, Ho, Wo, int(relu),
                         scale.data_ptr() if scale is not None else None, shift.data_ptr() if shift is not None else None,
                         dst.buf.data_ptr(), dst.cstride, dst.choff)
        op = Op()
        op.kind = OP_RESAMPLE
        op.u.resample = d
        self.ops.append(op)
        self.keep += [src.buf, dst.buf, scale, shift]
        self.records.append(("resample", dict(src=src, dst=dst, mode=mode, k=k, relu=relu, scale=scale, shift=shift)))

    def affine(self, src: View, dst: View, scale, shift, relu=True):
        """dst = act(src * scale[c] + shift[c]): BatchNorm(eval) -> ReLU ahead of a ResBlockA branch (_blocks.py:52-54)."""
        assert scale.shape == shift.shape == (src.channels,) and scale.dtype == shift.dtype == torch.float32
        self._resample(src, dst, 0, relu=relu, scale=scale, shift=shift)

    def leaky_relu(self, src: View, dst: View):
        """nn.LeakyReLU() with the default slope 0.01 (swinir.py:171)."""
        self._resample(src, dst, 0, relu=2)

    def winattn(self, qkv: View, biasT, heads, ws, shift, scale, out: View):
        """Shifted-window attention between the qkv and proj GEMMs of a SwinTransformerBlock (PSSR_OP_WINATTN)."""
        C = out.channels
        assert qkv.choff == 0 and qkv.channels == 3 * C and biasT.shape == (heads, ws * ws, ws * ws) and biasT.dtype == torch.float32
        d = WinAttnDesc(qkv.buf.data_ptr(), qkv.cstride, C, heads, qkv.B, qkv.H, qkv.W, ws, shift, float(scale), 0, biasT.data_ptr(),
                        out.buf.data_ptr(), out.cstride, out.choff)
        op = Op()
        op.kind = OP_WINATTN
        op.u.winattn = d
        self.ops.append(op)
        self.keep += [qkv.buf, biasT, out.buf]
        self.records.append(("winattn", dict(qkv=qkv, biasT=biasT, heads=heads, ws=ws, shift=shift, scale=float(scale), out=out)))

    def gather_channels(self, buf, choff, valid, dst: View):
        """dst[..., c] = buf[..., choff + c] for c < valid, zeros up to dst.channels: a torch.chunk piece at any channel offset."""
        assert dst.channels % 8 == 0 and valid <= dst.channels and choff + valid <= buf.shape[3] and buf.shape[:3] == dst.buf.shape[:3]
        d = ResampleDesc(buf.data_ptr(), buf.shape[3], choff, dst.channels, dst.B, dst.H, dst.W, 3, valid, dst.H, dst.W, 0, None, None,
                         dst.buf.data_ptr(), dst.cstride, dst.choff)
        op = Op()
        op.kind = OP_RESAMPLE
        op.u.resample = d
        self.ops.append(op)
        self.keep += [buf, dst.buf]
        self.records.append(("gather", dict(buf=buf, choff=choff, valid=valid, dst=dst)))

    def maxpool_k(self, src: View, dst: View, k):
        """F.max_pool2d(x, kernel_size=k) (PSP_Pooling, _blocks.py:85)."""
        assert (dst.H, dst.W) == (src.H // k, src.W // k)
        self._resample(src, dst, 1, k=k)

    def upsample_bilinear(self, src: View, dst: View):
        """F.interpolate(x, size=(dst.H, dst.W), mode="bilinear") (PSP_Pooling, _blocks.py:85)."""
        self._resample(src, dst, 2)

    def maxpool(self, src: View, dst: View):
        d = PoolDesc(src.buf.data_ptr(), src.cstride, src.choff, dst.buf.data_ptr(), dst.cstride, dst.choff, src.B, src.H,
                     src.W, src.channels)
        op = Op()
        op.kind = OP_MAXPOOL
        op.u.pool = d
        self.ops.append(op)
        self.keep += [src.buf, dst.buf]
        self.records.append(("maxpool", dict(src=src, dst=dst)))

    def tail(self, src: View, weight, bias, mul, add, out_f32=None, out_u8=None):
        cout = weight.shape[0]
        d = TailDesc(src.ptr(), src.cstride, src.channels, src.B, src.H, src.W, weight.data_ptr(), bias.data_ptr(), cout,
                     mul, add, out_f32.data_ptr() if out_f32 is not None else None,
                     out_u8.data_ptr() if out_u8 is not None else None)
        op = Op()
        op.kind = OP_TAIL
        op.u.tail = d
        self.ops.append(op)
        self.keep += [src.buf, weight, bias, out_f32, out_u8]
        self.records.append(("tail", dict(src=src, weight=weight, bias=bias, mul=mul, add=add, out_f32=out_f32, out_u8=out_u8)))

    def tailsum(self, z, r, bias, mul, add, out_f32=None, out_u8=None, layout=0):
        B, H, planes, W = z.shape            # z[b][y][s*9+t][x]  (layout 1: 48 window sums per LR pixel, include/pssr_b200.h)
        assert planes == (48 if layout == 1 else r * r * 9) and z.dtype == torch.float32 and z.is_contiguous()
        d = TailSumDesc(z.data_ptr(), B, H, W, r, float(bias), mul, add, layout, out_f32.data_ptr() if out_f32 is not None else None,
                        out_u8.data_ptr() if out_u8 is not None else None)
        op = Op()
        op.kind = OP_TAILSUM
        op.u.tailsum = d
        self.ops.append(op)
        self.keep += [z, out_f32, out_u8]
        self.records.append(("tailsum", dict(z=z, r=r, bias=float(bias), mul=mul, add=add, out_f32=out_f32, out_u8=out_u8, layout=layout)))

    @staticmethod
    def _lo_ptr(v: View, lo: View):
        """Compensated precision: a (hi, lo) pair shares one layout; returns the base pointer of the lo buffer (or None)."""
        if lo is None:
            return None
        assert lo.buf.shape == v.buf.shape and lo.buf.dtype == v.buf.dtype and (lo.choff, lo.channels) == (v.choff, v.channels)
        return lo.buf.data_ptr()

    def stem(self, x, in_scale, in_shift, patch, weight, bias, ln_w, ln_b, eps, out: View, out_lo: View = None):
        B, C, H, W = x.shape
        cout = weight.shape[0]
        d = StemDesc(x.data_ptr(), 1 if x.dtype == torch.uint8 else 0, B, C, H, W, in_scale.data_ptr(), in_shift.data_ptr(), patch, cout,
                     weight.data_ptr(), bias.data_ptr(), ln_w.data_ptr(), ln_b.data_ptr(), eps, 0, out.buf.data_ptr(), out.cstride, out.choff,
                     self._lo_ptr(out, out_lo))
        op = Op()
        op.kind = OP_STEM
        op.u.stem = d
        self.ops.append(op)
        self.keep += [x, in_scale, in_shift, weight, bias, ln_w, ln_b, out.buf, out_lo.buf if out_lo is not None else None]
        self.records.append(("stem", dict(x=x, in_scale=in_scale, in_shift=in_shift, patch=patch, weight=weight, bias=bias, ln_w=ln_w,
                                          ln_b=ln_b, eps=eps, out=out, out_lo=out_lo)))

    def layernorm(self, src: View, w, b, eps, out: View, s2d=1, src_lo: View = None, out_lo: View = None):
        d = LnDesc(src.buf.data_ptr(), src.cstride, src.choff, src.channels, src.B, src.H, src.W, s2d, w.data_ptr(), b.data_ptr(), eps, 0,
                   out.buf.data_ptr(), out.cstride, out.choff, self._lo_ptr(src, src_lo), self._lo_ptr(out, out_lo))
        op = Op()
        op.kind = OP_LAYERNORM
        op.u.ln = d
        self.ops.append(op)
        self.keep += [src.buf, w, b, out.buf] + [v.buf for v in (src_lo, out_lo) if v is not None]
        self.records.append(("ln", dict(src=src, w=w, b=b, eps=eps, out=out, s2d=s2d, src_lo=src_lo, out_lo=out_lo)))

    def dwconv_ln(self, src: View, dw_w, dw_b, ln_w, ln_b, eps, out: View, src_lo: View = None, out_lo: View = None):
        assert (src_lo is None) == (out_lo is None)
        d = DwLnDesc(src.buf.data_ptr(), src.cstride, src.choff, src.channels, src.B, src.H, src.W, 0, dw_w.data_ptr(), dw_b.data_ptr(),
                     ln_w.data_ptr(), ln_b.data_ptr(), eps, 0, out.buf.data_ptr(), out.cstride, out.choff, self._lo_ptr(src, src_lo),
                     self._lo_ptr(out, out_lo))
        op = Op()
        op.kind = OP_DWCONV_LN
        op.u.dwln = d
        self.ops.append(op)
        self.keep += [src.buf, dw_w, dw_b, ln_w, ln_b, out.buf] + [v.buf for v in (src_lo, out_lo) if v is not None]
        self.records.append(("dwln", dict(src=src, dw_w=dw_w, dw_b=dw_b, ln_w=ln_w, ln_b=ln_b, eps=eps, out=out, src_lo=src_lo, out_lo=out_lo)))

    def ese(self, src: View, fc_w, fc_b, gamma, gate_ws, out: View):
        d = EseDesc(src.buf.data_ptr(), src.cstride, src.channels, src.B, src.H, src.W, 0, fc_w.data_ptr(), fc_b.data_ptr(),
                    gamma.data_ptr() if gamma is not None else None, gate_ws.data_ptr(), out.buf.data_ptr(), out.cstride, out.choff)
        op = Op()
        op.kind = OP_ESE
        op.u.ese = d
        self.ops.append(op)
        self.keep += [src.buf, fc_w, fc_b, gamma, gate_ws, out.buf]
        self.records.append(("ese", dict(src=src, fc_w=fc_w, fc_b=fc_b, gamma=gamma, out=out)))

    # ---- lifecycle ---------------------------------------------------------------------
    def finalize(self):
        if DRY_RUN:   # tests/: build the op list on CPU tensors without creating the native plan
            return self
        arr = (Op * len(self.ops))(*self.ops)
        h = ctypes.c_void_p()
        # the plan belongs to the device its buffers live on: created and run with that device current
        self.device = _lib.same_device(*[t for t in self.keep if isinstance(t, torch.Tensor)])
        with _lib.on_device(self.device):
            _lib.check(_lib.lib().pssr_plan_create(arr, len(self.ops), self.dtype, ctypes.byref(h)), "pssr_plan_create")
        self.handle = h
        return self

    def run(self, first=None, count=None):
        if self.handle is None:
            raise RuntimeError("plan not finalized")
        with _lib.on_device(self.device):
            st = _lib.current_stream_ptr(self.device)
            if first is None:
                _lib.check(_lib.lib().pssr_plan_run(self.handle, st), "pssr_plan_run")
            else:
                _lib.check(_lib.lib().pssr_plan_run_range(self.handle, first, count, st), "pssr_plan_run_range")

    def __len__(self):
        return len(self.ops)

    def close(self):
        if self.handle is not None:
            _lib.lib().pssr_plan_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
