"""CPU tests (-m "not gpu"): the oracle against the committed golden vectors (produced by the real
reference, tests/golden/gen_golden.py), against Pillow itself, and against brute-force formulas for
the parity-unpinned third-party restatements."""
import os

import numpy as np
import pytest
import torch

from oracle import pipeline as OP
from oracle import thirdparty as TP
from oracle.models import resunet_forward
from oracle.pillow_resize import pillow_resize, resize_bilinear

G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("dtype,h,w,oh,ow", [(np.uint8, 512, 512, 128, 128), (np.uint16, 512, 512, 128, 128), (np.uint8, 256, 256, 32, 32),
                                             (np.uint16, 500, 500, 125, 125), (np.uint8, 96, 96, 32, 32), (np.uint16, 100, 260, 25, 65),
                                             (np.uint8, 64, 64, 64, 64)])
def test_resize_restatement_matches_pillow(dtype, h, w, oh, ow):
    rng = np.random.default_rng(h + w)
    img = rng.integers(0, 256 if dtype == np.uint8 else 65536, (2, h, w)).astype(dtype)
    assert np.array_equal(resize_bilinear(img, oh, ow), pillow_resize(img, oh, ow))


def test_gen_pair_golden():
    g = np.load(os.path.join(G, "gen_pair.npz"))
    for tag in ("u8_s4", "u16_s4", "u8_s8_pad", "u16_s2_frames"):
        hr_res, scale, has_frames = (int(v) for v in g[f"{tag}_meta"])
        stages = [("poisson", g[f"{tag}_poisson"], 0.8, 2), ("gaussian", g[f"{tag}_normal"])]
        hr, lr = OP.gen_pair(g[f"{tag}_in"], hr_res, scale, stages, n_frames=[3, 1] if has_frames else None)
        assert np.array_equal(hr, g[f"{tag}_hr"]), tag
        assert np.array_equal(lr, g[f"{tag}_lr"]), tag
    hr, lr = OP.gen_pair(g["none_in"], 96, 3, None)
    assert np.array_equal(hr, g["none_hr"]) and np.array_equal(lr, g["none_lr"])


def test_tiling_and_stitch_golden():
    g = np.load(os.path.join(G, "tiling_stitch.npz"))
    sheet = g["sheet"]
    for tag in "abc":
        size, stride, nf, slide, tx, ty, n_slices = (int(v) for v in g[f"meta_{tag}"])
        nf = None if nf < 0 else nf
        assert OP.n_tiles(sheet.shape[-2:], size, stride) == (tx, ty)
        tiles = np.stack([OP.sliding_window(sheet, size, stride, nf, n_slices, i, bool(slide)) for i in range(tx * ty * n_slices)])
        assert np.array_equal(tiles, g[f"tiles_{tag}"])
    for tag in ("p0", "p1", "p2", "p3"):
        n_rows, n_cols, T, ov, margin = (int(v) for v in g[f"{tag}_meta"])
        assert np.array_equal(OP.stitch_sheets(g[f"{tag}_tiles"], n_rows, n_cols, ov, margin)[0], g[f"{tag}_sheet"])


def test_normalize_golden():
    g = np.load(os.path.join(G, "normalize.npz"))
    a, b = OP.normalize_preds(g["hr"], g["hat"])
    assert np.array_equal(a, g["hr_norm"]) and np.array_equal(b, g["hat_norm"])


def test_net_golden():
    from pssr2_b200.models import ResUNet
    g = np.load(os.path.join(G, "net.npz"))
    for tag, kw in [("resunet_small", dict(hidden=[64, 128], scale=2, depth=1)),
                    ("resunet_5ch_s8", dict(channels=[5, 1], hidden=[64, 128], scale=8, depth=0))]:
        torch.manual_seed(1234)
        m = ResUNet(**kw).eval()     # same module tree => same parameter order => same seeded init as the reference
        gen = torch.Generator().manual_seed(1)
        for mod in m.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.running_mean.copy_(torch.randn(mod.running_mean.shape, generator=gen) * 0.1)
                mod.running_var.copy_(torch.rand(mod.running_var.shape, generator=gen) + 0.5)
        wsum = float(sum(p.double().sum() for p in m.state_dict().values() if p.is_floating_point()))
        if abs(wsum - float(g[f"{tag}_wsum"][0])) > 1e-6:
            pytest.skip("torch's seeded initialisation differs from the generator run; golden weights not reproducible here")
        y = resunet_forward(m.state_dict(), torch.as_tensor(g[f"{tag}_x"]))
        assert float((y - torch.as_tensor(g[f"{tag}_y"])).abs().max()) < 2e-4
        assert (OP.pred_array(y.numpy()) != g[f"{tag}_pred"]).mean() < 1e-3


VARIANT_CASES = [("resunet_atrous", "ResUNet", dict(hidden=[64, 128, 256], dilations=[[1, 3, 15], [1, 3], [1]], depth=1, scale=2)),
                 ("resunet_psp", "ResUNet", dict(hidden=[64, 128], pool_sizes=[1, 2, 4, 8], encoder_pool=True, depth=1, scale=4)),
                 ("resunet_a_default_dil", "ResUNet", dict(channels=[3, 1], hidden=[64, 128], dilations=[[1, 3, 15, 31], [1, 3, 15]], pool_sizes=[1, 2, 4, 8], depth=0, scale=2)),
                 ("rdresunet_a", "RDResUNet", dict(hidden=[128, 128], growth_rates=[32, 40, 64], ds_blocks=[False, True, False], ese_blocks=[False, True, True],
                                                   n_blocks=[2, 1, 2], rdnet_init=64, scale=2, depth=1, dilations=[[1], [1, 3]], pool_sizes=[1, 2, 4, 8]))]


def variant_model(tag, cls, kw, g):
    """The repo's module for a golden variant case with the generator script's seeded weights (tests/golden/gen_golden.py
    net_variant_cases); None when torch's seeded initialisation does not reproduce them here."""
    import pssr2_b200.models as M
    torch.manual_seed(4321)
    m = getattr(M, cls)(**kw).eval()
    gen = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for mod in m.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.running_mean.copy_(torch.randn(mod.running_mean.shape, generator=gen) * 0.1)
                mod.running_var.copy_(torch.rand(mod.running_var.shape, generator=gen) + 0.5)
                mod.weight.copy_(torch.rand(mod.weight.shape, generator=gen) * 0.4 + 0.8)
                mod.bias.copy_(torch.randn(mod.bias.shape, generator=gen) * 0.1)
        for n, p_ in m.named_parameters():
            if n.endswith("gamma"):
                p_.copy_(torch.rand(p_.shape, generator=gen) * 0.5 + 0.25)
    wsum = float(sum(p.double().sum() for p in m.state_dict().values() if p.is_floating_point()))
    return m if abs(wsum - float(g[f"{tag}_wsum"][0])) <= 1e-6 else None


def variant_oracle(cls, kw, sd, x):
    from oracle.models import rdresunet_forward
    if cls == "ResUNet":
        return resunet_forward(sd, x, dilations=kw.get("dilations"), pool_sizes=kw.get("pool_sizes"))
    return rdresunet_forward(sd, x, ds_blocks=kw["ds_blocks"], dilations=kw.get("dilations"), pool_sizes=kw.get("pool_sizes"))


@pytest.mark.parametrize("tag,cls,kw", VARIANT_CASES)
def test_net_variants_golden(tag, cls, kw):
    """Atrous blocks / PSP pooling (pssr/models/_blocks.py:43-92): the oracle restatement against the reference's own output."""
    g = np.load(os.path.join(G, "net_variants.npz"))
    m = variant_model(tag, cls, kw, g)
    if m is None:
        pytest.skip("torch's seeded initialisation differs from the generator run; golden weights not reproducible here")
    y = variant_oracle(cls, kw, m.state_dict(), torch.as_tensor(g[f"{tag}_x"]))
    assert float((y - torch.as_tensor(g[f"{tag}_y"])).abs().max()) < 2e-4


SWINIR_CASES = [("swinir_small", dict(image_size=32, depths=[2, 2], num_heads=[6, 6])),
                ("swinir_w4_s2", dict(image_size=64, depths=[3], num_heads=[4], embed_dim=64, scale=2, channels=[3, 1], window_size=4))]


def swinir_model(tag, kw, g):
    """The repo's SwinIR with the generator script's seeded weights (tests/golden/gen_golden.py swinir_cases), or None."""
    from pssr2_b200.models import SwinIR
    torch.manual_seed(777)
    m = SwinIR(**kw).eval()
    gen = torch.Generator().manual_seed(2)
    with torch.no_grad():
        for n, p_ in m.named_parameters():
            if "relative_position_bias_table" in n or n.endswith("bias"):
                p_.copy_(torch.randn(p_.shape, generator=gen) * 0.2)
    wsum = float(sum(p.double().sum() for p in m.state_dict().values() if p.is_floating_point()))
    return m if abs(wsum - float(g[f"{tag}_wsum"][0])) <= 1e-6 else None


@pytest.mark.parametrize("tag,kw", SWINIR_CASES)
def test_swinir_golden(tag, kw):
    """SwinIR (pssr/models/swinir.py:221-258): the oracle restatement against the reference's own output."""
    from oracle.models import swinir_forward
    g = np.load(os.path.join(G, "swinir.npz"))
    m = swinir_model(tag, kw, g)
    if m is None:
        pytest.skip("torch's seeded initialisation differs from the generator run; golden weights not reproducible here")
    y = swinir_forward(m.state_dict(), torch.as_tensor(g[f"{tag}_x"]))
    assert float((y - torch.as_tensor(g[f"{tag}_y"])).abs().max()) < 2e-4


def test_ssim_psnr_bruteforce():
    """parity-unpinned restatement of skimage: check against a direct per-window evaluation."""
    rng = np.random.default_rng(0)
    a = rng.integers(0, 256, (24, 31)).astype(np.uint8)
    b = np.clip(a.astype(int) + rng.integers(-30, 31, a.shape), 0, 255).astype(np.uint8)
    C1, C2 = (0.01 * 255) ** 2, (0.03 * 255) ** 2
    vals = []
    for y in range(3, a.shape[0] - 3):
        for x in range(3, a.shape[1] - 3):
            wa = a[y - 3:y + 4, x - 3:x + 4].astype(np.float64).ravel()
            wb = b[y - 3:y + 4, x - 3:x + 4].astype(np.float64).ravel()
            ux, uy = wa.mean(), wb.mean()
            vx, vy = wa.var(ddof=1), wb.var(ddof=1)
            vxy = ((wa - ux) * (wb - uy)).sum() / 48
            vals.append((2 * ux * uy + C1) * (2 * vxy + C2) / ((ux ** 2 + uy ** 2 + C1) * (vx + vy + C2)))
    assert abs(TP.structural_similarity(a, b, data_range=255) - np.mean(vals)) < 1e-10
    mse = ((a.astype(float) - b.astype(float)) ** 2).mean()
    assert abs(TP.peak_signal_noise_ratio(a, b, data_range=255) - 10 * np.log10(255 ** 2 / mse)) < 1e-12
    with pytest.raises(ValueError):
        TP.structural_similarity(a[:5], b[:5], data_range=255)


def test_random_noise_sp_semantics():
    rng = np.random.default_rng(1)
    img = rng.random((8, 8)).astype(np.float32)
    fl = rng.random(img.shape) <= 0.3
    sa = rng.random(img.shape) <= 0.5
    out = TP.random_noise(img, "s&p", 0.3, flipped=fl, salted=sa)
    assert out.dtype == np.float32
    assert np.all(out[fl & sa] == 1) and np.all(out[fl & ~sa] == 0) and np.array_equal(out[~fl], img[~fl])


def test_gen_pair_rotation_golden():
    """Training-time rot90 / flip (pssr/data.py:476-480) against the reference's own _gen_pair outputs."""
    g = np.load(os.path.join(G, "gen_pair_rot.npz"))
    for k in range(6):
        rot, f1, f2 = (int(v) for v in g[f"code_{k}"])
        axes = (1, 2) if f1 and f2 else (1 if f1 else 2)
        hr, lr = OP.gen_pair(g["in"], 64, 4, None, rotation=[bool(rot), axes])
        assert np.array_equal(hr, g[f"hr_{k}"]) and np.array_equal(lr, g[f"lr_{k}"]), k
