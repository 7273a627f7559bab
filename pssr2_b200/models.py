"""Drop-in ``ResUNet`` / ``RDResUNet`` modules (reference: pssr/models/resunet.py:7-99,
pssr/models/_blocks.py:6-41, pssr/models/rdresunet.py:8-133).

The module trees keep the reference's parameter / buffer names, so ``load_state_dict`` accepts
reference checkpoints unchanged (SURVEY.md §8a state-dict contract).  ``forward`` does not run
PyTorch ops: at first use for a given input shape it folds BatchNorm into the weights, packs them
K-major for the tcgen05 implicit-GEMM kernel and builds a plan of fused CUDA ops
(``pssr2_b200/plan.py`` -> ``libpssr_b200.so``); later calls only launch that plan.  There is no
CPU or eager fallback: a non-CUDA input raises.
"""
import torch
import torch.nn as nn

from .plan import ACT_NONE, ACT_RELU, Plan, View, ceil_div, pack_weight, permute_n


def _force_list(item):
    """pssr/util.py:220-226."""
    if type(item) is not list:
        try:
            return list(item)
        except Exception:
            return [item]
    return item


# ------------------------------------------------------------------------- module skeletons
class ResBlock(nn.Module):
    """Parameter container mirroring pssr/models/_blocks.py:20-41 (conv.{0,3,6,..} convs,
    conv.{1,4,7,..} BatchNorm2d, conv.{2,5,..} ReLU, respass 1x1)."""

    def __init__(self, in_channels, out_channels, depth, norm=True):
        super().__init__()
        layers = []
        n_layers = max(depth, 0) + 1
        for i in range(n_layers):
            layers.append(nn.Conv2d(in_channels if i == 0 else out_channels, out_channels, kernel_size=3, padding=1))
            if norm:
                layers.append(nn.BatchNorm2d(out_channels))
            if i + 1 < n_layers:
                layers.append(nn.ReLU(inplace=True))
        self.conv = nn.Sequential(*layers)
        self.respass = nn.Conv2d(in_channels, out_channels, kernel_size=1)
        self.depth = depth
        self.in_channels, self.out_channels = in_channels, out_channels

    def folded(self):
        """[(W*s, b*s+t)] per 3x3 conv with BatchNorm(eval) folded, plus the respass (W, b)."""
        out = []
        mods = list(self.conv)
        i = 0
        while i < len(mods):
            conv = mods[i]
            w, b = conv.weight.detach().float(), conv.bias.detach().float()
            i += 1
            if i < len(mods) and isinstance(mods[i], nn.BatchNorm2d):
                bn = mods[i]
                s = bn.weight.detach().float() / torch.sqrt(bn.running_var.float() + bn.eps)
                t = bn.bias.detach().float() - bn.running_mean.float() * s
                w, b = w * s.view(-1, 1, 1, 1), b * s + t
                i += 1
            if i < len(mods) and isinstance(mods[i], nn.ReLU):
                i += 1
            out.append((w, b))
        return out, (self.respass.weight.detach().float(), self.respass.bias.detach().float())


class Reconstruction(nn.Module):
    """Parameter container mirroring pssr/models/_blocks.py:6-18."""

    def __init__(self, in_channels, out_channels, hidden, scale=4):
        super().__init__()
        self.pre = nn.Conv2d(hidden + in_channels, scale ** 2 * hidden, kernel_size=3, padding=1)
        self.conv = nn.Conv2d(hidden, out_channels, kernel_size=3, padding=1)
        self.scale = scale


class _PlanModule(nn.Module):
    """Shared forward machinery: plan cache keyed by input geometry, invalidated whenever the
    parameters may have changed (load_state_dict, .to(), train())."""

    precision = "fp16"   # operand format of the tensor-core path: "fp16" or "bf16" (fp32 accumulate)
    fuse_tail = True     # fuse Reconstruction.conv into Reconstruction.pre's epilogue (single output channel)

    def __init__(self):
        super().__init__()
        self._plans = {}

    def invalidate(self):
        for st in self._plans.values():
            st["plan"].close()
        self._plans = {}

    def _apply(self, fn, *a, **k):
        self.invalidate()
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self.invalidate()
        return super().load_state_dict(*a, **k)

    def train(self, mode=True):
        if mode:
            self.invalidate()
        return super().train(mode)

    def _state(self, x):
        if not x.is_cuda:
            raise RuntimeError("pssr2_b200 models run on CUDA (sm_100a) tensors only; there is no CPU fallback")
        if self.training:
            raise RuntimeError("pssr2_b200 models implement the eval()/predict path only (call model.eval())")
        if x.dtype not in (torch.float32, torch.uint8):
            x = x.float()
        key = (tuple(x.shape), x.dtype, x.device.index, self.precision, self.fuse_tail)
        st = self._plans.get(key)
        if st is None:
            with torch.no_grad():
                st = self._build(x.shape, x.dtype, x.device)
            self._plans[key] = st
        return st, x

    @torch.no_grad()
    def forward(self, x):
        st, x = self._state(x)
        st["x"].copy_(x)
        st["plan"].run()
        return st["out"].clone()   # the plan's output buffer is reused by the next call

    @torch.no_grad()
    def forward_u8(self, x):
        """forward + `_pred_array` (pssr/predict.py:245-246) fused on device: returns the fp32 output
        and the uint8 [B,1,H,W] truncation of its centre channel."""
        st, x = self._state(x)
        st["x"].copy_(x)
        st["plan"].run()
        return st["out"], st["out_u8"]

    # helpers used by subclasses ---------------------------------------------------------
    @staticmethod
    def _emit_resblock(plan, blk, srcs, seg_spec, w0_parts_fn, wr_parts_fn, scratch, dst, shuffle, B, H, W):
        """Emits the convs of one ResBlock.
        srcs / seg_spec: views and (src, taps, cblocks) segments feeding conv0 (3x3) -- and, with taps
        forced to the 1x1 variant by ``wr_parts_fn``, the respass.  ``w0_parts_fn(w)`` / ``wr_parts_fn(w)``
        split a conv0 / respass weight into per-segment [Cout, Cin_seg, kh, kw] parts."""
        convs, (wr, br) = blk.folded()
        cout = blk.out_channels
        dev = wr.device
        n = len(convs)
        cb_out = ceil_div(cout, 64)
        cur = None
        for i, (w, b) in enumerate(convs):
            last = i + 1 == n
            if i == 0:
                in_srcs, in_segs, parts = list(srcs), list(seg_spec), w0_parts_fn(w)
            else:
                in_srcs, in_segs, parts = [cur], [(0, 9, cb_out)], [w]
            bias = b
            if last:
                # relu(conv_n(h) + respass(x)): the 1x1 residual is extra K blocks of the same GEMM
                r_parts, r_segs = wr_parts_fn(wr)
                base = len(in_srcs)
                for v in srcs:
                    in_srcs.append(v)
                in_segs += [(base + si, taps, cb) for (si, taps, cb) in r_segs]
                parts = parts + r_parts
                bias = b + br
                # merge duplicate source views (conv0 == last conv when depth == 0)
            wp = pack_weight(parts, plan.dtype, shuffle if last else 1)
            bp = permute_n(bias, shuffle if last else 1).contiguous()
            in_srcs, in_segs = _dedupe_sources(in_srcs, in_segs)
            out_view = dst if last else View(scratch[i % 2])
            plan.conv(in_srcs, in_segs, wp, bp, out_view, Ho=H, Wo=W, B=B, shuffle=shuffle if last else 1, act=ACT_RELU)
            for p in parts:
                plan.flops += 2 * p.numel() * B * H * W
            cur = out_view
        return dst


def _dedupe_sources(srcs, segs):
    """The C descriptor holds at most 3 distinct source views; identical views are merged."""
    uniq, remap = [], {}
    for i, v in enumerate(srcs):
        key = (v.buf.data_ptr(), v.choff, v.channels)
        for j, u in enumerate(uniq):
            if (u.buf.data_ptr(), u.choff, u.channels) == key:
                remap[i] = j
                break
        else:
            remap[i] = len(uniq)
            uniq.append(v)
    return uniq, [(remap[s], t, c) for (s, t, c) in segs]


def _im2col_parts(w):
    """[Cout, C, 3, 3] -> a 1x1 weight over the 64-channel im2col tensor (channel = c*9 + tap)."""
    co, c = w.shape[:2]
    return w.reshape(co, c * 9, 1, 1)


def _im2col_centre(w):
    """[Cout, C, 1, 1] respass weight -> 1x1 weight over the im2col tensor (centre tap = c*9 + 4)."""
    co, c = w.shape[:2]
    out = torch.zeros(co, c * 9, 1, 1, device=w.device)
    out[:, torch.arange(c, device=w.device) * 9 + 4, 0, 0] = w[:, :, 0, 0]
    return out


class ResUNet(_PlanModule):
    """Residual UNet + upscaling block (pssr/models/resunet.py:7-99), default (non-atrous) path."""

    def __init__(self, channels=1, hidden=[64, 128, 256, 512, 1024], scale=4, depth=3, dilations=None, pool_sizes=None,
                 encoder_pool=False):
        super().__init__()
        channels = _force_list(channels)
        channels = channels * 2 if len(channels) == 1 else channels
        if dilations or pool_sizes or encoder_pool:
            raise NotImplementedError(
                "pssr2_b200.ResUNet implements the default residual path; the atrous / PSP-pooling variants "
                "(ResBlockA, PSP_Pooling) are outside the accelerated hot path")
        hidden = list(hidden)
        self.norm = nn.BatchNorm2d(channels[0])
        self.encoder, self.decoder = nn.ModuleList(), nn.ModuleList()
        layers = [channels[0], *hidden]
        n_layers = len(layers) - 1
        for i in range(n_layers):
            self.encoder.append(ResBlock(layers[i], layers[i + 1], depth))
            if i + 1 < n_layers:
                self.decoder.append(ResBlock(layers[-i - 1] - int(layers[-i - 2] / 2), layers[-i - 2], depth))
        self.encoder_pool = None
        self.reconstruction_pool = None
        self.reconstruction = Reconstruction(channels[0], channels[1], hidden[0], scale)
        self.channels, self.hidden, self.scale = channels, hidden, scale

    def extra_repr(self):
        return (f"ResUNet with {self.reconstruction.scale}x upscaling\n{len(self.encoder)} residual decoder blocks with "
                f"{self.encoder[0].depth} hidden layers each\nPSP pooling disabled")

    # -------------------------------------------------------------------------------------
    def _build(self, shape, in_dtype, dev):
        B, C, H, W = shape
        hid, L, s = self.hidden, len(self.hidden), self.scale
        if C != self.channels[0]:
            raise ValueError(f"expected {self.channels[0]} input channels, got {C}")
        if H % (1 << (L - 1)) or W % (1 << (L - 1)):
            raise ValueError(f"input size {H}x{W} must be divisible by {1 << (L - 1)}")
        for i in range(1, L):
            if hid[i] % 4 or (hid[i] // 4) % 8 or hid[i - 1] % 8:
                raise NotImplementedError(f"hidden={hid}: channel counts must keep 8-channel alignment after pixel shuffle")
        plan = Plan(self.precision)
        dt = plan.tdtype
        z = lambda *sh: torch.zeros(*sh, dtype=dt, device=dev)
        x_in = torch.zeros(B, C, H, W, dtype=in_dtype, device=dev)

        # input normalisation + im2col of the few-channel input (resunet.py:66-70)
        bn = self.norm
        sc = (bn.weight.detach().float() / torch.sqrt(bn.running_var.float() + bn.eps)).contiguous()
        sh = (bn.bias.detach().float() - bn.running_mean.float() * sc).contiguous()
        im2col = z(B, H, W, 64)
        plan.prep(x_in, sc, sh, im2col)
        xcol = View(im2col, 0, 64)

        # level l lives at H/2^l; cat[l] = [pixel_shuffle(decoder input), encoder skip l]
        up = [hid[l + 1] // 4 for l in range(L - 1)]
        cat = [z(B, H >> l, W >> l, up[l] + hid[l]) for l in range(L - 1)]
        scratch_elems = max(B * (H >> l) * (W >> l) * hid[l] for l in range(L))
        sbuf = [torch.zeros(scratch_elems, dtype=dt, device=dev) for _ in range(2)]

        def scratch(l):
            n = B * (H >> l) * (W >> l) * hid[l]
            return [sb[:n].view(B, H >> l, W >> l, hid[l]) for sb in sbuf]

        cur = None  # pooled input of the current encoder level
        deepest = None
        for l in range(L):
            blk = self.encoder[l]
            h, w = H >> l, W >> l
            if l == 0:
                srcs, segs = [xcol], [(0, 1, 1)]
                w0f = lambda wt: [_im2col_parts(wt)]
                wrf = lambda wt: ([_im2col_centre(wt)], [(0, 1, 1)])
            else:
                cin = hid[l - 1]
                srcs, segs = [cur], [(0, 9, ceil_div(cin, 64))]
                w0f = lambda wt: [wt]
                wrf = (lambda cin: (lambda wt: ([wt], [(0, 1, ceil_div(cin, 64))])))(cin)
            if l + 1 < L:
                dst = View(cat[l], up[l], hid[l])
                self._emit_resblock(plan, blk, srcs, segs, w0f, wrf, scratch(l), dst, 1, B, h, w)
                pooled = z(B, h // 2, w // 2, hid[l])
                plan.maxpool(dst, View(pooled))
                cur = View(pooled)
            else:
                # deepest block: its output is pixel-shuffled straight into the first decoder's concat buffer
                dst = View(cat[l - 1], 0, up[l - 1])
                self._emit_resblock(plan, blk, srcs, segs, w0f, wrf, scratch(l), dst, 2, B, h, w)
        # decoder (resunet.py:81-85): block j works at level l = L-2-j on cat[l]
        final = z(B, H, W, hid[0])
        for j in range(L - 1):
            l = L - 2 - j
            blk = self.decoder[j]
            h, w = H >> l, W >> l
            cin = up[l] + hid[l]
            srcs, segs = [View(cat[l], 0, cin)], [(0, 9, ceil_div(cin, 64))]
            w0f = lambda wt: [wt]
            wrf = (lambda cin: (lambda wt: ([wt], [(0, 1, ceil_div(cin, 64))])))(cin)
            if l > 0:
                dst, shf = View(cat[l - 1], 0, up[l - 1]), 2
            else:
                dst, shf = View(final), 1
            self._emit_resblock(plan, blk, srcs, segs, w0f, wrf, scratch(l), dst, shf, B, h, w)

        # Reconstruction (resunet.py:90-95, _blocks.py:15-18): cat([x, xnorm]) -> pre -> relu -> shuffle(s) -> conv
        rec = self.reconstruction
        wp = rec.pre.weight.detach().float()
        bp = rec.pre.bias.detach().float()
        parts = [wp[:, :hid[0]], _im2col_parts(wp[:, hid[0]:])]
        wpk = pack_weight(parts, plan.dtype, s)
        wc = rec.conv.weight.detach().float()
        bc = rec.conv.bias.detach().float().contiguous()
        cout = wc.shape[0]
        out = torch.empty(B, cout, H * s, W * s, dtype=torch.float32, device=dev)
        out_u8 = torch.empty(B, 1, H * s, W * s, dtype=torch.uint8, device=dev)
        srcs, segs = [View(final), xcol], [(0, 9, ceil_div(hid[0], 64)), (1, 1, 1)]
        plan.flops += 2 * wp.numel() * B * H * W + 2 * wc.numel() * B * H * s * W * s
        if cout == 1 and hid[0] % 32 == 0 and W >= 1 and self.fuse_tail:
            # fused tail: relu(pre) is reduced against the 3x3 tail weights inside the conv epilogue (fp32), the
            # scale^2*hidden-channel HR map is never written; PSSR_OP_TAILSUM gathers the 9 taps (see include/pssr_b200.h)
            tw = wc[0].permute(1, 2, 0).reshape(9, hid[0]).contiguous()      # [tap][c]
            zbuf = torch.zeros(B, s * s * 9, H, W, dtype=torch.float32, device=dev)
            plan.conv(srcs, segs, wpk, permute_n(bp, s).contiguous(), None, Ho=H, Wo=W, B=B, shuffle=s, act=ACT_RELU,
                      tail_weight=tw, tail_z=zbuf)
            plan.tailsum(zbuf, s, float(bc[0]), 128.0, 128.0, out, out_u8)    # x*128+128 (resunet.py:95)
        else:
            ps_out = z(B, H * s, W * s, hid[0])
            plan.conv(srcs, segs, wpk, permute_n(bp, s).contiguous(), View(ps_out), Ho=H, Wo=W, B=B, shuffle=s, act=ACT_RELU)
            plan.tail(View(ps_out), wc.permute(0, 2, 3, 1).contiguous(), bc, 128.0, 128.0, out, out_u8)
        plan.finalize()
        return {"plan": plan, "x": x_in, "out": out, "out_u8": out_u8}
