"""fp32 CPU restatement of the reference network forwards, driven by a plain state_dict.

TEST INFRASTRUCTURE ONLY.  Follows
  ResUNet.forward       pssr/models/resunet.py:65-96
  ResBlock.forward      pssr/models/_blocks.py:39-41   (conv/BN/ReLU stack :23-33)
  Reconstruction        pssr/models/_blocks.py:15-18
  RDResUNet.forward     pssr/models/rdresunet.py:104-130
  RDNet.forward         pssr/models/_rdnet.py:95-104, DenseStage :133-138,
                        DenseBlock :168-175, Block/BlockESE :181-203
Pinned against the reference modules themselves (imported via oracle/refshim.py) in
tests/test_oracle_vs_reference.py and through tests/golden/net_*.npz.

``emulate`` ("bf16" | "fp16" | None) additionally models the arithmetic of the CUDA plan
(BN folded into 16-bit weights, 16-bit activations in HBM, fp32 accumulation and fp32
epilogues); it is used only to derive the tolerances written in the GPU parity tests.
"""
import torch
import torch.nn.functional as F

_EMU = {"bf16": torch.bfloat16, "fp16": torch.float16, None: None}


def _q(t, emulate):
    dt = _EMU[emulate]
    return t if dt is None else t.to(dt).to(torch.float32)


def _bn_fold(sd, prefix, eps=1e-5):
    """BatchNorm2d(eval) as y = s*x + t  (torch default eps, pssr/models/_blocks.py:30)."""
    s = sd[prefix + ".weight"] / torch.sqrt(sd[prefix + ".running_var"] + eps)
    t = sd[prefix + ".bias"] - sd[prefix + ".running_mean"] * s
    return s, t


def _n_convs(sd, prefix):
    n = 0
    while f"{prefix}.conv.{3 * n}.weight" in sd:
        n += 1
    return n


def resblock(sd, prefix, x, emulate=None):
    """relu( [conv3x3 -> BN -> ReLU]*depth -> conv3x3 -> BN  +  conv1x1(x) )  (_blocks.py:20-41)."""
    n = _n_convs(sd, prefix)
    if emulate is None:
        h = x
        for i in range(n):
            h = F.conv2d(h, sd[f"{prefix}.conv.{3 * i}.weight"], sd[f"{prefix}.conv.{3 * i}.bias"], padding=1)
            s, t = _bn_fold(sd, f"{prefix}.conv.{3 * i + 1}")
            h = h * s.view(1, -1, 1, 1) + t.view(1, -1, 1, 1)
            if i + 1 < n:
                h = F.relu(h)
        r = F.conv2d(x, sd[f"{prefix}.respass.weight"], sd[f"{prefix}.respass.bias"])
        return F.relu(h + r)
    # emulation of the plan: folded weights, 16-bit operands, fp32 accumulate
    h = x  # already quantised by the caller
    for i in range(n):
        w = sd[f"{prefix}.conv.{3 * i}.weight"]
        b = sd[f"{prefix}.conv.{3 * i}.bias"]
        s, t = _bn_fold(sd, f"{prefix}.conv.{3 * i + 1}")
        wf = _q(w * s.view(-1, 1, 1, 1), emulate)
        bf = b * s + t
        acc = F.conv2d(h, wf, None, padding=1) + bf.view(1, -1, 1, 1)
        if i + 1 < n:
            h = _q(F.relu(acc), emulate)
        else:
            wr = _q(sd[f"{prefix}.respass.weight"], emulate)
            acc = acc + F.conv2d(x, wr, sd[f"{prefix}.respass.bias"])
            h = _q(F.relu(acc), emulate)
    return h


def resblock_a(sd, prefix, x):
    """ResBlockA (_blocks.py:43-68): relu( sum_d [BN -> ReLU -> conv3x3(dilation d, padding "same")]*(depth+1) (x)  +  conv1x1(x) );
    the dilation of branch j is recovered from the state_dict by the caller (it is not stored): ``sd["__dilations__"][prefix]``."""
    dils = sd["__dilations__"][prefix]
    total = F.conv2d(x, sd[f"{prefix}.respass.weight"], sd[f"{prefix}.respass.bias"])
    for j, d in enumerate(dils):
        h = x
        i = 0
        while f"{prefix}.dilations.{j}.{3 * i + 2}.weight" in sd:
            s, t = _bn_fold(sd, f"{prefix}.dilations.{j}.{3 * i}")
            h = F.relu(h * s.view(1, -1, 1, 1) + t.view(1, -1, 1, 1))
            h = F.conv2d(h, sd[f"{prefix}.dilations.{j}.{3 * i + 2}.weight"], sd[f"{prefix}.dilations.{j}.{3 * i + 2}.bias"], padding=d, dilation=d)
            i += 1
        total = total + h
    return F.relu(total)


def any_resblock(sd, prefix, x, emulate=None):
    if f"{prefix}.dilations.0.0.weight" in sd:
        return resblock_a(sd, prefix, x)
    return resblock(sd, prefix, x, emulate)


def psp_pooling(sd, prefix, x, sizes):
    """PSP_Pooling.forward (_blocks.py:80-92): channel chunks -> max_pool(size) -> bilinear back to the input size -> conv1x1 + BN + ReLU;
    concat -> conv1x1 + BN -> ReLU."""
    size = x.shape[-2:]
    outs = []
    for i, chunk in enumerate(torch.chunk(x, len(sizes), 1)):
        c = F.interpolate(F.max_pool2d(chunk, kernel_size=sizes[i]), size=size, mode="bilinear")
        c = F.conv2d(c, sd[f"{prefix}.convs.{i}.0.weight"], sd[f"{prefix}.convs.{i}.0.bias"])
        s, t = _bn_fold(sd, f"{prefix}.convs.{i}.1")
        outs.append(F.relu(c * s.view(1, -1, 1, 1) + t.view(1, -1, 1, 1)))
    x = F.conv2d(torch.cat(outs, 1), sd[f"{prefix}.conv_out.weight"], sd[f"{prefix}.conv_out.bias"])
    s, t = _bn_fold(sd, f"{prefix}.norm_out")
    return F.relu(x * s.view(1, -1, 1, 1) + t.view(1, -1, 1, 1))


def reconstruction(sd, x, scale, emulate=None):
    """relu(pre(x)) -> pixel_shuffle(scale) -> conv   (_blocks.py:15-18)."""
    wp, bp = sd["reconstruction.pre.weight"], sd["reconstruction.pre.bias"]
    wc, bc = sd["reconstruction.conv.weight"], sd["reconstruction.conv.bias"]
    h = F.relu(F.conv2d(x, _q(wp, emulate), bp, padding=1))
    h = _q(h, emulate)
    h = F.pixel_shuffle(h, scale)
    # the tail conv runs on CUDA cores with fp32 weights in the plan
    return F.conv2d(h, wc, bc, padding=1)


def _input_norm(sd, x):
    x = x / 128 - 1                                    # resunet.py:66
    if "norm.weight" in sd:                            # resunet.py:67-68 (BN, eval)
        s, t = _bn_fold(sd, "norm")
        x = x * s.view(1, -1, 1, 1) + t.view(1, -1, 1, 1)
    return x


def _sd_float(sd, dilations=None, prefixes=()):
    out = {k: v.float() for k, v in sd.items() if torch.is_tensor(v) and v.is_floating_point()}
    if dilations:
        out["__dilations__"] = dict(zip(prefixes, dilations))
    return out


def resunet_forward(sd, x, scale=None, emulate=None, dilations=None, pool_sizes=None):
    """ResUNet.forward (resunet.py:65-96).  ``dilations`` (one list per hidden level, as the constructor takes them) selects the
    atrous blocks, ``pool_sizes`` the PSP pooling layers the state_dict holds (reconstruction_pool, encoder_pool)."""
    n_lv = 0
    while f"encoder.{n_lv}.respass.weight" in sd:
        n_lv += 1
    pre = [f"encoder.{i}" for i in range(n_lv)] + [f"decoder.{i}" for i in range(n_lv - 1)]
    dl = (list(dilations) + [dilations[-i - 1] for i in range(n_lv - 1)]) if dilations else None      # resunet.py:56-58
    sd = _sd_float(sd, dl, pre)
    if scale is None:
        hidden0 = sd["reconstruction.conv.weight"].shape[1]
        scale = int(round((sd["reconstruction.pre.weight"].shape[0] / hidden0) ** 0.5))
    n_enc = 0
    while f"encoder.{n_enc}.respass.weight" in sd:
        n_enc += 1
    x = _q(_input_norm(sd, x.float()), emulate)
    skips = [x]
    for i in range(n_enc):
        x = any_resblock(sd, f"encoder.{i}", x, emulate)
        if i + 1 < n_enc:
            skips.append(x)
            x = F.max_pool2d(x, 2)
    if "encoder_pool.conv_out.weight" in sd:             # resunet.py:78-79
        x = psp_pooling(sd, "encoder_pool", x, pool_sizes)
    for i in range(n_enc - 1):
        x = F.pixel_shuffle(x, 2)
        x = torch.cat([x, skips.pop()], 1)
        x = any_resblock(sd, f"decoder.{i}", x, emulate)
    if "reconstruction_pool.conv_out.weight" in sd:      # resunet.py:87-88
        x = psp_pooling(sd, "reconstruction_pool", x, pool_sizes)
    x = torch.cat([x, skips.pop()], 1)
    x = reconstruction(sd, x, scale, emulate)
    return x * 128 + 128                               # resunet.py:95


# ------------------------------------------------------------------------------ RDNet
def _ln2d(x, w, b, eps=1e-6):
    x = x.permute(0, 2, 3, 1)
    x = F.layer_norm(x, (x.shape[-1],), w, b, eps)
    return x.permute(0, 3, 1, 2)


def _rd_block(sd, p, x):
    """Block / BlockESE (_rdnet.py:177-206) followed by layer-scale gamma (:172-174)."""
    L = p + ".layers.layers"
    c = x.shape[1]
    h = F.conv2d(x, sd[L + ".0.weight"], sd[L + ".0.bias"], padding=3, groups=c)
    h = _ln2d(h, sd[L + ".1.weight"], sd[L + ".1.bias"])
    h = F.conv2d(h, sd[L + ".2.weight"], sd[L + ".2.bias"])
    h = F.gelu(h)
    h = F.conv2d(h, sd[L + ".4.weight"], sd[L + ".4.bias"])
    if L + ".5.fc.weight" in sd:
        se = h.mean((2, 3), keepdim=True)
        se = F.conv2d(se, sd[L + ".5.fc.weight"], sd[L + ".5.fc.bias"])
        h = h * (F.relu6(se + 3.0) / 6.0)
    if p + ".gamma" in sd:
        h = h * sd[p + ".gamma"].view(1, -1, 1, 1)
    return h


def rdnet_forward(sd, x, ds_blocks, prefix="encoder"):
    """RDNet.forward (_rdnet.py:95-104): stem, dense stages, skips before each downsample."""
    P = prefix
    w = sd[P + ".stem.stem.0.weight"]
    x = F.conv2d(x, w, sd[P + ".stem.stem.0.bias"], stride=w.shape[-1])
    x = _ln2d(x, sd[P + ".stem.stem.1.weight"], sd[P + ".stem.stem.1.bias"])
    skips = []
    for i, ds in enumerate(ds_blocks):
        if ds:
            skips.append(x)
        S = f"{P}.dense_stages.{i}"
        stage_idx = 0
        if f"{S}.0.weight" in sd:  # transition: LayerNorm2d + conv (k = stride = 1 or 2)
            x = _ln2d(x, sd[f"{S}.0.weight"], sd[f"{S}.0.bias"])
            wt = sd[f"{S}.1.weight"]
            x = F.conv2d(x, wt, sd[f"{S}.1.bias"], stride=wt.shape[-1])
            stage_idx = 2
        feats = [x]
        j = 0
        while f"{S}.{stage_idx}.dense_block{j}.layers.layers.0.weight" in sd:
            feats.append(_rd_block(sd, f"{S}.{stage_idx}.dense_block{j}", torch.cat(feats, 1)))
            j += 1
        x = torch.cat(feats, 1)
    return skips + [x]


def rdresunet_forward(sd, x, ds_blocks=(False, True, True, False, False, False, True), scale=None,
                      patch_size=2, dilations=None, pool_sizes=None):
    """RDResUNet.forward (rdresunet.py:104-130); ``dilations`` / ``pool_sizes`` as in ``resunet_forward``."""
    nd = 0
    while f"decoder.{nd}.respass.weight" in sd:
        nd += 1
    sd = _sd_float(sd, dilations, [f"decoder.{i}" for i in range(nd)])
    hidden_last = sd["reconstruction.conv.weight"].shape[1]
    if scale is None:
        scale = int(round((sd["reconstruction.pre.weight"].shape[0] / hidden_last) ** 0.5))
    x = _input_norm(sd, x.float())
    skips = [x] + rdnet_forward(sd, x, ds_blocks)
    n_dec = 0
    while f"decoder.{n_dec}.respass.weight" in sd:
        n_dec += 1
    ratios = [1] + [2] * (n_dec - 1) + [patch_size]
    if "encoder_pool.conv_out.weight" in sd:             # rdresunet.py:111-112
        skips[-1] = psp_pooling(sd, "encoder_pool", skips[-1], pool_sizes)
    for i in range(n_dec):
        x = torch.cat([x, skips.pop()], 1) if i != 0 else skips.pop()
        x = any_resblock(sd, f"decoder.{i}", x)
        x = F.pixel_shuffle(x, ratios[i + 1])
    if "reconstruction_pool.conv_out.weight" in sd:      # rdresunet.py:120-121
        x = psp_pooling(sd, "reconstruction_pool", x, pool_sizes)
    x = torch.cat([x, skips.pop()], 1)
    x = reconstruction(sd, x, scale)
    return x * 128 + 128


# ------------------------------------------------------------------------------ SwinIR
def _swin_windows(t, ws):
    """window_partition (swinir.py:722-736): [B, H, W, C] -> [B * nW, ws * ws, C]."""
    B, H, W, C = t.shape
    return t.view(B, H // ws, ws, W // ws, ws, C).permute(0, 1, 3, 2, 4, 5).reshape(-1, ws * ws, C)


def _swin_mask(H, W, ws, shift):
    """SwinTransformerBlock.calculate_mask (swinir.py:320-341)."""
    img = torch.zeros(1, H, W, 1)
    cnt = 0
    for hs in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
        for wsl in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
            img[:, hs, wsl, :] = cnt
            cnt += 1
    mw = _swin_windows(img, ws).view(-1, ws * ws)
    m = mw.unsqueeze(1) - mw.unsqueeze(2)
    return m.masked_fill(m != 0, -100.0).masked_fill(m == 0, 0.0)


def swinir_forward(sd, x):
    """SwinIR.forward (pssr/models/swinir.py:221-258) for the default configuration family: upsampler "pixelshuffle", resi_connection
    "1conv", patch_norm, no absolute position embedding, patch_size 1; depths, heads, window size and scale are read off the
    state_dict.  Blocks: SwinTransformerBlock.forward :343-388, WindowAttention.forward :563-592, RSTB.forward :449-450."""
    sd = {k: v.float() if v.is_floating_point() else v for k, v in sd.items()}
    x = x.float()
    H0, W0 = x.shape[2:]
    tab0 = sd["layers.0.residual_group.blocks.0.attn.relative_position_bias_table"]
    ws = (int(round(tab0.shape[0] ** 0.5)) + 1) // 2
    x = F.pad(x, (0, (ws - W0 % ws) % ws, 0, (ws - H0 % ws) % ws), "reflect")          # check_image_size :201-206
    B, _, H, W = x.shape
    if min(H, W) <= ws:
        raise NotImplementedError("maps no larger than one window change the block geometry (swinir.py:298-301)")
    f0 = F.conv2d(x, sd["conv_first.weight"], sd["conv_first.bias"], padding=1)
    C = f0.shape[1]
    ln = lambda t, p: F.layer_norm(t, (C,), sd[p + ".weight"], sd[p + ".bias"], 1e-5)
    t = f0.flatten(2).transpose(1, 2)
    t = ln(t, "patch_embed.norm")
    i = 0
    while f"layers.{i}.conv.weight" in sd:
        res = t
        j = 0
        while f"layers.{i}.residual_group.blocks.{j}.norm1.weight" in sd:
            P = f"layers.{i}.residual_group.blocks.{j}"
            shift = 0 if j % 2 == 0 else ws // 2
            tab, idx = sd[P + ".attn.relative_position_bias_table"], sd[P + ".attn.relative_position_index"]
            nh = tab.shape[1]
            y = ln(t, P + ".norm1").view(B, H, W, C)
            if shift:
                y = torch.roll(y, shifts=(-shift, -shift), dims=(1, 2))
            win = _swin_windows(y, ws)
            qkv = F.linear(win, sd[P + ".attn.qkv.weight"], sd[P + ".attn.qkv.bias"]).reshape(-1, ws * ws, 3, nh, C // nh).permute(2, 0, 3, 1, 4)
            attn = (qkv[0] * (C // nh) ** -0.5) @ qkv[1].transpose(-2, -1)
            attn = attn + tab[idx.view(-1)].view(ws * ws, ws * ws, -1).permute(2, 0, 1).unsqueeze(0)
            if shift:
                m = _swin_mask(H, W, ws, shift)
                attn = (attn.view(B, m.shape[0], nh, ws * ws, ws * ws) + m.unsqueeze(1).unsqueeze(0)).view(-1, nh, ws * ws, ws * ws)
            y = (attn.softmax(-1) @ qkv[2]).transpose(1, 2).reshape(-1, ws * ws, C)
            y = F.linear(y, sd[P + ".attn.proj.weight"], sd[P + ".attn.proj.bias"])
            y = y.view(B, H // ws, W // ws, ws, ws, C).permute(0, 1, 3, 2, 4, 5).reshape(B, H, W, C)       # window_reverse
            if shift:
                y = torch.roll(y, shifts=(shift, shift), dims=(1, 2))
            t = t + y.view(B, H * W, C)
            h = F.gelu(F.linear(ln(t, P + ".norm2"), sd[P + ".mlp.fc1.weight"], sd[P + ".mlp.fc1.bias"]))
            t = t + F.linear(h, sd[P + ".mlp.fc2.weight"], sd[P + ".mlp.fc2.bias"])
            j += 1
        img = t.transpose(1, 2).reshape(B, C, H, W)
        img = F.conv2d(img, sd[f"layers.{i}.conv.weight"], sd[f"layers.{i}.conv.bias"], padding=1)
        t = img.flatten(2).transpose(1, 2) + res
        i += 1
    t = ln(t, "norm")
    f = F.conv2d(t.transpose(1, 2).reshape(B, C, H, W), sd["conv_after_body.weight"], sd["conv_after_body.bias"], padding=1) + f0
    f = F.leaky_relu(F.conv2d(f, sd["conv_before_upsample.0.weight"], sd["conv_before_upsample.0.bias"], padding=1), 0.01)
    k, scale = 0, 1
    while f"upsample.{k}.weight" in sd:
        w = sd[f"upsample.{k}.weight"]
        r = int(round((w.shape[0] / w.shape[1]) ** 0.5))
        f = F.pixel_shuffle(F.conv2d(f, w, sd[f"upsample.{k}.bias"], padding=1), r)
        scale *= r
        k += 2
    out = F.conv2d(f, sd["conv_last.weight"], sd["conv_last.bias"], padding=1)
    return out[:, :, :H0 * scale, :W0 * scale]
