"""Drop-in ``SwinIR`` (reference: pssr/models/swinir.py:16-268 and the blocks below it) behind the same plan API as the UNets.

The module tree keeps the reference's parameter / buffer names and registration order, so ``load_state_dict`` accepts reference
checkpoints unchanged.  ``forward`` launches a plan: every ``nn.Linear`` (qkv, proj, fc1, fc2) is a 1x1 GEMM and every convolution
a 3x3 implicit GEMM on the tcgen05 kernels over the NHWC token map (residual adds ride in the GEMM epilogues, GELU in fc1's),
LayerNorm is the ``ln`` kernel, the shifted-window attention between qkv and proj is ``PSSR_OP_WINATTN`` (csrc/swin.cu) and the last
convolution + `_pred_array` is the ``tail`` op.  Supported: the default family -- upsampler "pixelshuffle", resi_connection "1conv",
patch_size 1, no absolute position embedding, scale 2^n; inputs whose size is a multiple of the window (the reference reflect-pads
others, swinir.py:201-206).
"""
import math

import torch
import torch.nn as nn

from .models import _PlanModule, _force_list
from .plan import ACT_GELU, ACT_NONE, Plan, View, ceil_div, pack_weight


class _Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.act = nn.GELU()
        self.fc2 = nn.Linear(hidden, dim)


class WindowAttention(nn.Module):
    """Parameter container for pssr/models/swinir.py:523-592."""

    def __init__(self, dim, window_size, num_heads, qkv_bias=True, qk_scale=None):
        super().__init__()
        self.dim, self.window_size, self.num_heads = dim, window_size, num_heads
        head_dim = dim // num_heads
        self.scale = qk_scale or head_dim ** -0.5
        ws = window_size
        self.relative_position_bias_table = nn.Parameter(torch.zeros((2 * ws - 1) * (2 * ws - 1), num_heads))
        coords = torch.stack(torch.meshgrid([torch.arange(ws), torch.arange(ws)], indexing="ij")).flatten(1)
        rel = (coords[:, :, None] - coords[:, None, :]).permute(1, 2, 0).contiguous()
        rel[:, :, 0] += ws - 1
        rel[:, :, 1] += ws - 1
        rel[:, :, 0] *= 2 * ws - 1
        self.register_buffer("relative_position_index", rel.sum(-1))
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)
        nn.init.trunc_normal_(self.relative_position_bias_table, std=.02)


def _calculate_mask(H, W, ws, shift):
    """pssr/models/swinir.py:320-341 (kept as a buffer for state_dict compatibility; the kernel derives it from the regions)."""
    img = torch.zeros((1, H, W, 1))
    cnt = 0
    for hs in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
        for wsl in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
            img[:, hs, wsl, :] = cnt
            cnt += 1
    mw = img.view(1, H // ws, ws, W // ws, ws, 1).permute(0, 1, 3, 2, 4, 5).contiguous().view(-1, ws * ws)
    m = mw.unsqueeze(1) - mw.unsqueeze(2)
    return m.masked_fill(m != 0, float(-100.0)).masked_fill(m == 0, float(0.0))


class SwinTransformerBlock(nn.Module):
    """Parameter container for pssr/models/swinir.py:276-388."""

    def __init__(self, dim, input_resolution, num_heads, window_size, shift_size, mlp_ratio, qkv_bias, qk_scale):
        super().__init__()
        self.dim, self.input_resolution, self.num_heads = dim, input_resolution, num_heads
        self.window_size, self.shift_size = window_size, shift_size
        if min(input_resolution) <= window_size:
            self.shift_size = 0
            self.window_size = min(input_resolution)
        if not 0 <= self.shift_size < self.window_size:
            raise ValueError(f"shift_size must between 0 and window_size. Given values are {shift_size} and {window_size}.")
        self.norm1 = nn.LayerNorm(dim)
        self.attn = WindowAttention(dim, self.window_size, num_heads, qkv_bias, qk_scale)
        self.norm2 = nn.LayerNorm(dim)
        self.mlp = _Mlp(dim, int(dim * mlp_ratio))
        mask = _calculate_mask(input_resolution[0], input_resolution[1], self.window_size, self.shift_size) if self.shift_size > 0 else None
        self.register_buffer("attn_mask", mask)


class BasicLayer(nn.Module):
    def __init__(self, dim, input_resolution, depth, num_heads, window_size, mlp_ratio, qkv_bias, qk_scale):
        super().__init__()
        self.blocks = nn.ModuleList([SwinTransformerBlock(dim, input_resolution, num_heads, window_size, 0 if i % 2 == 0 else window_size // 2,
                                                          mlp_ratio, qkv_bias, qk_scale) for i in range(depth)])


class RSTB(nn.Module):
    """Parameter container for pssr/models/swinir.py:390-450 (resi_connection "1conv")."""

    def __init__(self, dim, input_resolution, depth, num_heads, window_size, mlp_ratio, qkv_bias, qk_scale):
        super().__init__()
        self.residual_group = BasicLayer(dim, input_resolution, depth, num_heads, window_size, mlp_ratio, qkv_bias, qk_scale)
        self.conv = nn.Conv2d(dim, dim, 3, 1, 1)


class _PatchEmbed(nn.Module):
    def __init__(self, dim, norm):
        super().__init__()
        self.norm = nn.LayerNorm(dim) if norm else None


class SwinIR(_PlanModule):
    r"""SwinIR as detailed in Liang et al., 2021 (pssr/models/swinir.py:16-268); same constructor arguments as the reference."""
    precision = "fp16"

    def __init__(self, image_size=128, channels=1, scale=4, embed_dim=96, mlp_ratio=2, depths=[4, 4, 4, 4], num_heads=[6, 6, 6, 6], window_size=8,
                 patch_size=1, upsampler="pixelshuffle", qkv_bias=True, qk_scale=None, drop_rate=0, attn_drop_rate=0, drop_path_rate=0.1,
                 norm_layer=nn.LayerNorm, ape=False, patch_norm=True, use_checkpoint=False, resi_connection="1conv"):
        super().__init__()
        if len(depths) != len(num_heads):
            raise ValueError(f"Lengths of depths and num_heads must be equal. Given lengths are {len(depths)} and {len(num_heads)}.")
        if upsampler != "pixelshuffle" or resi_connection != "1conv" or patch_size != 1 or ape or norm_layer is not nn.LayerNorm:
            raise NotImplementedError("pssr2_b200.SwinIR implements the default family: upsampler='pixelshuffle', resi_connection='1conv', "
                                      "patch_size=1, ape=False, norm_layer=nn.LayerNorm")
        if scale & (scale - 1):
            raise NotImplementedError(f"scale {scale}: the plan implements the 2^n pixel-shuffle stack (swinir.py:702-705)")
        channels = _force_list(channels)
        channels = channels * 2 if len(channels) == 1 else channels
        num_feat = 64
        self.upscale, self.upsampler, self.window_size = scale, upsampler, window_size
        self.conv_first = nn.Conv2d(channels[0], embed_dim, 3, 1, 1)
        self.num_layers, self.embed_dim, self.mlp_ratio = len(depths), embed_dim, mlp_ratio
        res = (image_size // patch_size, image_size // patch_size)
        self.patch_embed = _PatchEmbed(embed_dim, patch_norm)
        self.layers = nn.ModuleList([RSTB(embed_dim, res, depths[i], num_heads[i], window_size, mlp_ratio, qkv_bias, qk_scale) for i in range(len(depths))])
        self.norm = nn.LayerNorm(embed_dim)
        self.conv_after_body = nn.Conv2d(embed_dim, embed_dim, 3, 1, 1)
        self.conv_before_upsample = nn.Sequential(nn.Conv2d(embed_dim, num_feat, 3, 1, 1), nn.LeakyReLU(inplace=True))
        up = []
        for _ in range(int(math.log(scale, 2))):
            up += [nn.Conv2d(num_feat, 4 * num_feat, 3, 1, 1), nn.PixelShuffle(2)]
        self.upsample = nn.Sequential(*up)
        self.conv_last = nn.Conv2d(num_feat, channels[1], 3, 1, 1)
        self.channels = channels
        self.apply(self._init_weights)

    @staticmethod
    def _init_weights(m):
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    def extra_repr(self):
        return f"SwinIR with {self.upscale}x upscaling\n{self.num_layers} Swin Transformer blocks with embedding size {self.embed_dim}"

    # -----------------------------------------------------------------------------------------
    def _build(self, shape, in_dtype, dev):
        B, Cin, H, W = shape
        C, ws = self.embed_dim, self.window_size
        if Cin != self.channels[0]:
            raise ValueError(f"expected {self.channels[0]} input channels, got {Cin}")
        if H % ws or W % ws:
            raise NotImplementedError(f"input size {H}x{W}: the plan needs multiples of the window size {ws} (the reference reflect-pads, swinir.py:201-206)")
        if min(H, W) <= ws:
            raise NotImplementedError("maps no larger than one window change the block geometry (swinir.py:298-301)")
        if C % 32 or Cin > 64:
            raise NotImplementedError(f"embed_dim {C} must be a multiple of 32 and the input have at most 64 channels")
        plan = Plan("fp16" if self.precision == "fp16c" else self.precision)
        dt = plan.tdtype
        z = lambda *sh: torch.zeros(*sh, dtype=dt, device=dev)
        f32 = lambda t: t.detach().float().contiguous()
        x_in = torch.zeros(B, Cin, H, W, dtype=in_dtype, device=dev)
        # (x - mean) * img_range with mean 0, range 1 (swinir.py:91-96,225-226): the plan's input op computes (x/128 - 1) * s + t
        xn = z(B, H, W, ceil_div(Cin, 8) * 8)
        plan.prep(x_in, torch.full((Cin,), 128.0, device=dev), torch.full((Cin,), 128.0, device=dev), xn, centre_only=True)

        def conv3(src, cin, mod, out, **kw):
            w = f32(mod.weight)
            plan.flops += 2 * w.numel() * B * src.H * src.W
            plan.conv([src], [(0, 9, ceil_div(cin, 64))], pack_weight([w], plan.dtype, kw.get("shuffle", 1)),
                      _perm(f32(mod.bias), kw.get("shuffle", 1)), out, Ho=src.H, Wo=src.W, B=B, **kw)

        def linear(src, mod, out, **kw):
            w = f32(mod.weight)
            n, k = w.shape
            if n % 32:
                raise NotImplementedError(f"linear layer width {n} must be a multiple of 32")
            plan.flops += 2 * w.numel() * B * H * W
            bias = f32(mod.bias) if mod.bias is not None else torch.zeros(n, device=dev)
            plan.conv([src], [(0, 1, ceil_div(k, 64))], pack_weight([w.view(n, k, 1, 1)], plan.dtype), bias, out, Ho=H, Wo=W, B=B, **kw)

        def _perm(b, r):
            from .plan import permute_n
            return permute_n(b, r).contiguous()

        f0 = z(B, H, W, C)
        conv3(View(xn, 0, Cin), Cin, self.conv_first, View(f0))
        X = [z(B, H, W, C), z(B, H, W, C)]
        T = [z(B, H, W, C), z(B, H, W, C)]
        Y, A = z(B, H, W, C), z(B, H, W, C)
        hidden = int(C * self.mlp_ratio)
        QKV, Hb = z(B, H, W, 3 * C), z(B, H, W, hidden)
        pe = self.patch_embed.norm
        if pe is not None:
            plan.layernorm(View(f0), f32(pe.weight), f32(pe.bias), pe.eps, View(X[0]))
            cur = X[0]
        else:
            cur = f0
        xi = 0
        for layer in self.layers:
            res = cur
            for blk in layer.residual_group.blocks:
                at = blk.attn
                N = blk.window_size ** 2
                biasT = at.relative_position_bias_table.detach().float()[at.relative_position_index.view(-1)].view(N, N, -1).permute(2, 1, 0).contiguous()
                plan.layernorm(View(cur), f32(blk.norm1.weight), f32(blk.norm1.bias), blk.norm1.eps, View(Y))
                linear(View(Y), at.qkv, View(QKV))
                plan.winattn(View(QKV), biasT, at.num_heads, blk.window_size, blk.shift_size, at.scale, View(A))
                linear(View(A), at.proj, View(T[1]), resid=View(cur))                      # x = shortcut + attn
                plan.layernorm(View(T[1]), f32(blk.norm2.weight), f32(blk.norm2.bias), blk.norm2.eps, View(Y))
                linear(View(Y), blk.mlp.fc1, View(Hb), act=ACT_GELU)
                linear(View(Hb), blk.mlp.fc2, View(T[0]), resid=View(T[1]))                 # x = x + mlp(norm2(x))
                cur = T[0]
            # RSTB: conv(residual_group(x)) + x (swinir.py:449-450); the two X buffers alternate as the stage residual
            nxt = X[1 - xi] if res is X[xi] else X[xi]
            conv3(View(cur), C, layer.conv, View(nxt), resid=View(res))
            cur = nxt
            xi = 0 if cur is X[0] else 1
        plan.layernorm(View(cur), f32(self.norm.weight), f32(self.norm.bias), self.norm.eps, View(Y))
        body = z(B, H, W, C)
        conv3(View(Y), C, self.conv_after_body, View(body), resid=View(f0))
        nf = self.conv_before_upsample[0].weight.shape[0]
        u = z(B, H, W, nf)
        conv3(View(body), C, self.conv_before_upsample[0], View(u), act=ACT_NONE)
        plan.leaky_relu(View(u), View(u))
        h, w = H, W
        for m in self.upsample:
            if isinstance(m, nn.Conv2d):
                nu = z(B, 2 * h, 2 * w, nf)
                conv3(View(u), nf, m, View(nu), shuffle=2)
                u, h, w = nu, 2 * h, 2 * w
        cout = self.conv_last.weight.shape[0]
        out = torch.empty(B, cout, h, w, dtype=torch.float32, device=dev)
        out_u8 = torch.empty(B, 1, h, w, dtype=torch.uint8, device=dev)
        wl = f32(self.conv_last.weight)
        plan.flops += 2 * wl.numel() * B * h * w
        plan.tail(View(u), wl.permute(0, 2, 3, 1).contiguous(), f32(self.conv_last.bias), 1.0, 0.0, out, out_u8)     # x / img_range + mean
        plan.finalize()
        return {"plan": plan, "x": x_in, "out": out, "out_u8": out_u8}
