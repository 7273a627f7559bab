"""CPU interpreter for the python-level records of a ``pssr2_b200.plan.Plan`` (test infrastructure).

It executes the op list with plain PyTorch fp32 ops, reading the SAME packed weights and writing the
SAME activation buffers the CUDA kernels would, so that weight packing, BatchNorm folding, K-segment
scheduling, concat offsets and the pixel-shuffle permutation are checked on a CPU-only box."""
import torch
import torch.nn.functional as F


def _view_nchw(v, pad_to=None):
    t = v.buf[..., v.choff:v.choff + v.channels].float().permute(0, 3, 1, 2)
    if pad_to is not None and pad_to > v.channels:
        t = F.pad(t, (0, 0, 0, 0, 0, pad_to - v.channels))
    return t


def _store_pair(o, lo, y):
    """y [B,H,W,C] fp32 -> o (16-bit) and, for the compensated precision, what that rounding dropped -> lo."""
    c = y.shape[-1]
    o.buf[..., o.choff:o.choff + c] = y.to(o.buf.dtype)
    if lo is not None:
        lo.buf[..., lo.choff:lo.choff + c] = (y - o.buf[..., o.choff:o.choff + c].float()).to(lo.buf.dtype)


def _load_pair(v, lo):
    x = _view_nchw(v)
    return x if lo is None else x + _view_nchw(lo)


def run_records(plan):
    for kind, r in plan.records:
        if kind == "prep":
            x = r["x"].float()
            B, C, H, W = x.shape
            xn = (x / 128 - 1) * r["scale"].view(1, -1, 1, 1) + r["shift"].view(1, -1, 1, 1)
            lo = r.get("im2col_lo")
            cols = xn if r.get("centre_only") else F.unfold(xn, 3, padding=1).view(B, C * 9, H, W)
            cols = cols.permute(0, 2, 3, 1)
            nc = cols.shape[-1]
            r["im2col"].zero_()
            r["im2col"][..., :nc] = cols.to(r["im2col"].dtype)
            if lo is not None:          # compensated precision: what the 16-bit rounding dropped
                lo.zero_()
                lo[..., :nc] = (cols - r["im2col"][..., :nc].float()).to(lo.dtype)
        elif kind == "cast8":
            s, d = r["src"], r["dst"]
            y = s.buf[..., s.choff:s.choff + s.channels].float() * r["scale"]
            d.buf[..., d.choff:d.choff + s.channels] = y.clamp(-57344.0, 57344.0).to(torch.float8_e5m2).view(torch.uint8)
        elif kind == "maxpool":
            s, d = r["src"], r["dst"]
            d.buf[..., d.choff:d.choff + s.channels] = F.max_pool2d(_view_nchw(s), 2).permute(0, 2, 3, 1).to(d.buf.dtype)
        elif kind == "tail":
            s = r["src"]
            w = r["weight"].permute(0, 3, 1, 2)  # [Cout][3][3][C] -> [Cout][C][3][3]
            y = (F.conv2d(_view_nchw(s), w, r["bias"], padding=1)) * r["mul"] + r["add"]
            if r["out_f32"] is not None:
                r["out_f32"].copy_(y)
            if r["out_u8"] is not None:
                c = y.shape[1] // 2
                r["out_u8"].copy_(y[:, c:c + 1].clamp(0, 255).to(torch.uint8))
        elif kind == "conv":
            W = r["weight"].float()
            W8 = r["weight8"].view(torch.float8_e5m2).float() if r.get("weight8") is not None else None
            n = W.shape[0]
            acc = None
            k0 = k8 = 0
            for sg in r["segs"]:
                si, taps, cb = sg[:3]
                v = r["srcs"][si]
                kw = taps * cb * 64
                if len(sg) > 3 and sg[3] == 1:     # e5m2 source x e5m2 weights (fp32 accumulate)
                    wseg = W8[:, k8:k8 + kw].reshape(n, taps, cb * 64)
                    k8 += kw
                    x = v.buf[..., v.choff:v.choff + v.channels].view(torch.float8_e5m2).float().permute(0, 3, 1, 2)
                    x = F.pad(x, (0, 0, 0, 0, 0, cb * 64 - v.channels))
                else:
                    wseg = W[:, k0:k0 + kw].reshape(n, taps, cb * 64)
                    k0 += kw
                    x = _view_nchw(v, cb * 64)
                dil = sg[4] if len(sg) > 4 else 1
                if taps == 9:
                    y = F.conv2d(x, wseg.permute(0, 2, 1).reshape(n, cb * 64, 3, 3), padding=dil, dilation=dil)
                elif taps == 1:
                    y = F.conv2d(x, wseg.permute(0, 2, 1).reshape(n, cb * 64, 1, 1))
                else:
                    y = F.conv2d(x, wseg.permute(0, 2, 1).reshape(n, cb * 64, 2, 2), stride=2)
                acc = y if acc is None else acc + y
            assert k0 == W.shape[1]
            acc = acc + r["bias"].view(1, -1, 1, 1)
            if r.get("resid") is not None:
                acc = acc + r["resid_scale"] * _view_nchw(r["resid"])[:, :acc.shape[1]]
            if r["act"] == 1:
                acc = F.relu(acc)
            elif r["act"] == 2:
                acc = F.gelu(acc)
            if r["out_scale"] is not None:
                acc = acc * r["out_scale"].view(1, -1, 1, 1)
            acc = acc[:, :r["n_valid"]]
            rr = r["shuffle"]
            if rr > 1:  # stored order n = (i*r+j)*cps + cc
                B, N, H, Wd = acc.shape
                cps = N // (rr * rr)
                acc = acc.view(B, rr, rr, cps, H, Wd).permute(0, 3, 4, 1, 5, 2).reshape(B, cps, H * rr, Wd * rr)
            if r.get("tail_z") is not None:
                # acc is the pixel-shuffled post-activation map [B, C', H*r, W*r] (fp32): per-tap 1x1 projection, stored
                # per LR row and sub-pixel:  z[b][y][s*9+t][x]
                tw = r["tail_weight"]                                   # [9][C']
                if not (r.get("tail_flags", 0) & 1):
                    # single-pass tail: the activation goes back to TMEM as 16-bit and the tap weights are a 16-bit operand tile;
                    # the compensated tail (PSSR_TAIL_COMP) carries hi + lo of both, i.e. fp32 to first order
                    acc = acc.to(r["weight"].dtype).float()
                    tw = tw.to(r["weight"].dtype).float()
                zt = torch.einsum("bchw,tc->bthw", acc, tw)             # [B, 9, H*r, W*r]
                B_, _, Hh, Wh = zt.shape
                zt = zt.view(B_, 9, Hh // rr, rr, Wh // rr, rr).permute(0, 3, 5, 1, 2, 4).reshape(B_, rr * rr * 9, Hh // rr, Wh // rr)
                if r.get("tail_layout", 0) == 1:
                    # PSSR_TAIL_WINDOW48: zHR[b][t][4y+i'][4x+j'] feeds the output at (i'-dy, j'-dx) relative to the LR pixel;
                    # plane e*24 + (oi+1)*4 + (oj+1-2e), e = j' // 2
                    assert rr == 4
                    zh = zt.view(B_, rr, rr, 9, Hh // rr, Wh // rr)             # [b][i'][j'][t][y][x]
                    z48 = torch.zeros(B_, Hh // rr, 48, Wh // rr)
                    for i_ in range(4):
                        for j_ in range(4):
                            for t in range(9):
                                dy, dx = t // 3 - 1, t % 3 - 1
                                e = j_ // 2
                                z48[:, :, e * 24 + (i_ - dy + 1) * 4 + (j_ - dx + 1 - 2 * e), :] += zh[:, i_, j_, t]
                    r["tail_z"].copy_(z48)
                    continue
                r["tail_z"].copy_(zt.permute(0, 2, 1, 3))
                continue
            o = r["out"]
            if o is not None:
                o.buf[..., o.choff:o.choff + acc.shape[1]] = acc.permute(0, 2, 3, 1).to(o.buf.dtype)
            ol = r.get("out_lo")
            if ol is not None:
                hi = o.buf[..., o.choff:o.choff + acc.shape[1]].float()
                ol.buf[..., ol.choff:ol.choff + acc.shape[1]] = (acc.permute(0, 2, 3, 1) - hi).to(ol.buf.dtype)
            if r["out_f32"] is not None:
                r["out_f32"][..., :acc.shape[1]] = acc.permute(0, 2, 3, 1)
        elif kind == "resample":
            v, o = r["src"], r["dst"]
            x = _view_nchw(v)
            if r["mode"] == 0:
                if r["scale"] is not None:
                    x = x * r["scale"].view(1, -1, 1, 1) + r["shift"].view(1, -1, 1, 1)
                y = F.relu(x) if r["relu"] == 1 else (F.leaky_relu(x, 0.01) if r["relu"] == 2 else x)
            elif r["mode"] == 1:
                y = F.max_pool2d(x, r["k"])
            else:
                y = F.interpolate(x, size=(o.H, o.W), mode="bilinear")
            o.buf[..., o.choff:o.choff + y.shape[1]] = y.permute(0, 2, 3, 1).to(o.buf.dtype)
        elif kind == "gather":
            o = r["dst"]
            o.buf[..., o.choff:o.choff + o.channels] = 0
            o.buf[..., o.choff:o.choff + r["valid"]] = r["buf"][..., r["choff"]:r["choff"] + r["valid"]]
        elif kind == "winattn":
            v, o = r["qkv"], r["out"]
            ws, sh, nh = r["ws"], r["shift"], r["heads"]
            t = v.buf[..., :v.channels].float()                       # [B, H, W, 3C]
            B_, H, W, C3 = t.shape
            C = C3 // 3
            t = torch.roll(t, shifts=(-sh, -sh), dims=(1, 2)) if sh else t
            win = t.view(B_, H // ws, ws, W // ws, ws, C3).permute(0, 1, 3, 2, 4, 5).reshape(-1, ws * ws, 3, nh, C // nh).permute(2, 0, 3, 1, 4)
            q, k, vv = win[0] * r["scale"], win[1], win[2]
            attn = q @ k.transpose(-2, -1) + r["biasT"].transpose(1, 2).unsqueeze(0)
            if sh:
                img = torch.zeros(1, H, W, 1)
                cnt = 0
                for hs in (slice(0, -ws), slice(-ws, -sh), slice(-sh, None)):
                    for wsl in (slice(0, -ws), slice(-ws, -sh), slice(-sh, None)):
                        img[:, hs, wsl, :] = cnt
                        cnt += 1
                mw = img.view(1, H // ws, ws, W // ws, ws, 1).permute(0, 1, 3, 2, 4, 5).reshape(-1, ws * ws)
                mask = (mw.unsqueeze(1) - mw.unsqueeze(2) != 0).float() * -100.0        # [nW, N, N]
                nW = mask.shape[0]
                attn = (attn.view(B_, nW, nh, ws * ws, ws * ws) + mask.unsqueeze(1).unsqueeze(0)).view(-1, nh, ws * ws, ws * ws)
            y = (attn.softmax(-1) @ vv).transpose(1, 2).reshape(B_, H // ws, W // ws, ws, ws, C).permute(0, 1, 3, 2, 4, 5).reshape(B_, H, W, C)
            y = torch.roll(y, shifts=(sh, sh), dims=(1, 2)) if sh else y
            o.buf[..., o.choff:o.choff + C] = y.to(o.buf.dtype)
        elif kind == "stem":
            x = r["x"].float()
            xn = (x / 128 - 1) * r["in_scale"].view(1, -1, 1, 1) + r["in_shift"].view(1, -1, 1, 1)
            pch = r["patch"]
            w = r["weight"].view(r["weight"].shape[0], x.shape[1], pch, pch)
            y = F.conv2d(xn, w, r["bias"], stride=pch).permute(0, 2, 3, 1)
            y = F.layer_norm(y, (y.shape[-1],), r["ln_w"], r["ln_b"], r["eps"])
            _store_pair(r["out"], r.get("out_lo"), y)
        elif kind == "ln":
            v, o = r["src"], r["out"]
            y = _load_pair(v, r.get("src_lo")).permute(0, 2, 3, 1)
            y = F.layer_norm(y, (y.shape[-1],), r["w"], r["b"], r["eps"])
            if r["s2d"] == 2:
                B_, H, W, C = y.shape
                y = y.view(B_, H // 2, 2, W // 2, 2, C).permute(0, 1, 3, 2, 4, 5).reshape(B_, H // 2, W // 2, 4 * C)
            _store_pair(o, r.get("out_lo"), y)
        elif kind == "dwln":
            v, o = r["src"], r["out"]
            x = _load_pair(v, r.get("src_lo"))
            C = x.shape[1]
            w = r["dw_w"].t().reshape(C, 1, 7, 7)
            y = F.conv2d(x, w, r["dw_b"], padding=3, groups=C).permute(0, 2, 3, 1)
            y = F.layer_norm(y, (C,), r["ln_w"], r["ln_b"], r["eps"])
            _store_pair(o, r.get("out_lo"), y)
        elif kind == "ese":
            v, o = r["src"], r["out"]
            x = _view_nchw(v)
            C = x.shape[1]
            se = x.mean((2, 3), keepdim=True)
            se = F.conv2d(se, r["fc_w"].view(C, C, 1, 1), r["fc_b"])
            y = x * (F.relu6(se + 3.0) / 6.0)
            if r["gamma"] is not None:
                y = y * r["gamma"].view(1, -1, 1, 1)
            o.buf[..., o.choff:o.choff + C] = y.permute(0, 2, 3, 1).to(o.buf.dtype)
        elif kind == "tailsum" and r.get("layout", 0) == 1:
            z48, rr = r["z"], r["r"]
            B_, H, _, W = z48.shape
            canvas = torch.zeros(B_, H * 4 + 2, W * 4 + 2)
            for e in range(2):
                for oi in range(6):
                    for ojl in range(4):
                        canvas[:, oi:oi + 4 * H:4, ojl + 2 * e:ojl + 2 * e + 4 * W:4] += z48[:, :, e * 24 + oi * 4 + ojl, :]
            y = (canvas[:, None, 1:-1, 1:-1] + r["bias"]) * r["mul"] + r["add"]
            if r["out_f32"] is not None:
                r["out_f32"].copy_(y)
            if r["out_u8"] is not None:
                r["out_u8"].copy_(y.clamp(0, 255).to(torch.uint8))
        elif kind == "tailsum":
            z, rr = r["z"].permute(0, 2, 1, 3).contiguous(), r["r"]
            B_, _, H, W = z.shape
            zh = z.view(B_, rr, rr, 9, H, W).permute(0, 3, 4, 1, 5, 2).reshape(B_, 9, H * rr, W * rr)   # zHR[b][t][Y][X]
            zp = F.pad(zh, (1, 1, 1, 1))
            acc = torch.full((B_, 1, H * rr, W * rr), r["bias"])
            for t in range(9):
                dy, dx = t // 3 - 1, t % 3 - 1
                acc[:, 0] += zp[:, t, 1 + dy:1 + dy + H * rr, 1 + dx:1 + dx + W * rr]
            y = acc * r["mul"] + r["add"]
            if r["out_f32"] is not None:
                r["out_f32"].copy_(y)
            if r["out_u8"] is not None:
                r["out_u8"].copy_(y.clamp(0, 255).to(torch.uint8))
        else:
            raise ValueError(kind)
