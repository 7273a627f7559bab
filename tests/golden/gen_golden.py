"""Generates the committed golden vectors by running the UNMODIFIED reference (/root/reference/pssr,
imported through oracle/refshim.py) in the build container.  The reference cannot travel to the GPU
box, these small fixtures do.  Run:  python tests/golden/gen_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.refshim import import_reference  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
pssr = import_reference()
from pssr import crappifiers as RC, data as RD, util as RU  # noqa: E402
from pssr.models import ResUNet, RDResUNet  # noqa: E402
from pssr.predict import _pred_array  # noqa: E402


class Recorder:
    """Wraps np.random.poisson / normal so the reference's own draws can be replayed on the device."""

    def __enter__(self):
        self.draws = []
        self.p, self.n = np.random.poisson, np.random.normal

        def poisson(lam, *a, **k):
            y = self.p(lam, *a, **k)
            self.draws.append(("poisson", np.asarray(y)))
            return y

        def normal(loc=0.0, scale=1.0, size=None):
            g = self.n(loc, scale, size)
            self.draws.append(("normal", np.asarray(g)))
            return g

        np.random.poisson, np.random.normal = poisson, normal
        return self

    def __exit__(self, *a):
        np.random.poisson, np.random.normal = self.p, self.n


def gen_pair_cases():
    rng = np.random.default_rng(42)
    out = {}
    for tag, dtype, hr_res, scale, shape in [("u8_s4", np.uint8, 128, 4, (2, 128, 128)), ("u16_s4", np.uint16, 128, 4, (1, 128, 128)),
                                             ("u8_s8_pad", np.uint8, 128, 8, (1, 120, 124)), ("u16_s2_frames", np.uint16, 64, 2, (5, 64, 64))]:
        hr = rng.poisson(90, shape).clip(0, 255).astype(dtype)
        crap = RC.MultiCrappifier(RC.Poisson(intensity=0.8, gain=2), RC.AdditiveGaussian(intensity=9, gain=-1))
        n_frames = [3, 1] if "frames" in tag else None
        np.random.seed(7)
        with Recorder() as rec:
            h, l = RD._gen_pair(hr, hr_res, scale, False, crap, None, n_frames)
        out[f"{tag}_in"] = hr
        out[f"{tag}_hr"] = h.numpy()
        out[f"{tag}_lr"] = l.numpy()
        out[f"{tag}_poisson"] = rec.draws[0][1].astype(np.int64)
        out[f"{tag}_normal"] = rec.draws[1][1].astype(np.float64)
        out[f"{tag}_meta"] = np.array([hr_res, scale, 0 if n_frames is None else 1])
    # crappifier=None (no round / clip)
    hr = rng.integers(0, 256, (1, 96, 96)).astype(np.uint8)
    h, l = RD._gen_pair(hr, 96, 3, False, None, None, None)
    out["none_in"], out["none_hr"], out["none_lr"] = hr, h.numpy(), l.numpy()
    np.savez_compressed(os.path.join(OUT, "gen_pair.npz"), **out)


def gen_pair_rotation_cases():
    """Training-time augmentation (pssr/data.py:476-480): every (rot90, flip axes) combination through the reference's _gen_pair,
    crappifier=None, on a tile that needs reflect padding."""
    rng = np.random.default_rng(43)
    out = {}
    hr = rng.integers(0, 60000, (2, 44, 60)).astype(np.uint16)
    out["in"] = hr
    k = 0
    for rot in (False, True):
        for axes in (1, 2, (1, 2)):
            h, l = RD._gen_pair(hr, 64, 4, [rot, axes], None, None, None)
            out[f"hr_{k}"], out[f"lr_{k}"] = h.numpy(), l.numpy()
            out[f"code_{k}"] = np.array([int(rot), int(1 in (axes if isinstance(axes, tuple) else (axes,))),
                                         int(2 in (axes if isinstance(axes, tuple) else (axes,)))])
            k += 1
    np.savez_compressed(os.path.join(OUT, "gen_pair_rot.npz"), **out)


def tiling_stitch_cases():
    rng = np.random.default_rng(3)
    sheet = rng.integers(0, 256, (4, 150, 209)).astype(np.uint8)
    out = {"sheet": sheet}
    for tag, size, stride, nf, slide in [("a", 64, 48, None, False), ("b", 32, 32, 2, False), ("c", 50, 37, 3, True)]:
        tx, ty = RD._n_tiles(sheet, size, stride)
        n_slices = 1 if nf is None else ((sheet.shape[0] - nf + 1) if slide else sheet.shape[0] // nf)
        tiles = np.stack([RD._sliding_window(sheet, size, stride, nf, n_slices, i, slide) for i in range(tx * ty * n_slices)])
        out[f"tiles_{tag}"] = tiles
        out[f"meta_{tag}"] = np.array([size, stride, -1 if nf is None else nf, int(slide), tx, ty, n_slices])
    for tag, n_rows, n_cols, T, ov, margin in [("p0", 3, 4, 32, 8, 0), ("p1", 3, 3, 32, 8, 4), ("p2", 2, 5, 32, 8, 6), ("p3", 4, 4, 16, 10, 2)]:
        tiles = rng.integers(0, 256, (n_rows * n_cols, T, T)).astype(np.uint8)
        out[f"{tag}_tiles"] = tiles
        out[f"{tag}_sheet"] = np.asarray(RU._patch_images(tiles, n_cols, n_rows, ov, margin), dtype=np.uint8)
        out[f"{tag}_meta"] = np.array([n_rows, n_cols, T, ov, margin])
    out["val_idx_a"] = np.array(RD._get_val_idx([2, 3, 1, 4], 0.5, 0, [3, 2, 4, 1]))
    out["val_idx_b"] = np.array(RD._get_val_idx([1] * 10, 0.1, 0))
    np.savez_compressed(os.path.join(OUT, "tiling_stitch.npz"), **out)


def normalize_cases():
    rng = np.random.default_rng(5)
    base = rng.poisson(90, (3, 1, 96, 96)).clip(0, 255)
    hr = base.astype(np.uint8)
    hat = np.clip(base * 0.8 + 20 + rng.normal(0, 6, base.shape), 0, 255).astype(np.uint8)
    a, b = RU.normalize_preds(hr, hat)
    np.savez_compressed(os.path.join(OUT, "normalize.npz"), hr=hr, hat=hat, hr_norm=a, hat_norm=b)


def net_cases():
    out = {}
    rng = np.random.default_rng(9)
    for tag, cls, kw, shape in [("resunet_small", ResUNet, dict(hidden=[64, 128], scale=2, depth=1), (2, 1, 32, 32)),
                                ("resunet_5ch_s8", ResUNet, dict(channels=[5, 1], hidden=[64, 128], scale=8, depth=0), (1, 5, 16, 16))]:
        torch.manual_seed(1234)
        m = cls(**kw).eval()
        g = torch.Generator().manual_seed(1)
        for mod in m.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.running_mean.copy_(torch.randn(mod.running_mean.shape, generator=g) * 0.1)
                mod.running_var.copy_(torch.rand(mod.running_var.shape, generator=g) + 0.5)
        x = torch.tensor(rng.integers(0, 256, shape).astype(np.float32))
        with torch.no_grad():
            y = m(x)
        out[f"{tag}_x"] = x.numpy()
        out[f"{tag}_y"] = y.numpy()
        out[f"{tag}_pred"] = _pred_array(y)
        out[f"{tag}_wsum"] = np.array([float(sum(p.double().sum() for p in m.state_dict().values() if p.is_floating_point()))])
    np.savez_compressed(os.path.join(OUT, "net.npz"), **out)


VARIANT_CASES = [("resunet_atrous", "ResUNet", dict(hidden=[64, 128, 256], dilations=[[1, 3, 15], [1, 3], [1]], depth=1, scale=2), (2, 1, 32, 32)),
                 ("resunet_psp", "ResUNet", dict(hidden=[64, 128], pool_sizes=[1, 2, 4, 8], encoder_pool=True, depth=1, scale=4), (1, 1, 32, 32)),
                 ("resunet_a_default_dil", "ResUNet", dict(channels=[3, 1], hidden=[64, 128], dilations=[[1, 3, 15, 31], [1, 3, 15]], pool_sizes=[1, 2, 4, 8], depth=0, scale=2), (1, 3, 64, 64)),
                 ("rdresunet_a", "RDResUNet", dict(hidden=[128, 128], growth_rates=[32, 40, 64], ds_blocks=[False, True, False], ese_blocks=[False, True, True],
                                                   n_blocks=[2, 1, 2], rdnet_init=64, scale=2, depth=1, dilations=[[1], [1, 3]], pool_sizes=[1, 2, 4, 8]), (2, 1, 32, 32))]


def randomise_variant(m):
    """Non-trivial BatchNorm statistics / affine parameters and RDNet layer scales (the default gamma = 1e-6 hides the dense blocks)."""
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for mod in m.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.running_mean.copy_(torch.randn(mod.running_mean.shape, generator=g) * 0.1)
                mod.running_var.copy_(torch.rand(mod.running_var.shape, generator=g) + 0.5)
                mod.weight.copy_(torch.rand(mod.weight.shape, generator=g) * 0.4 + 0.8)
                mod.bias.copy_(torch.randn(mod.bias.shape, generator=g) * 0.1)
        for n, p_ in m.named_parameters():
            if n.endswith("gamma"):
                p_.copy_(torch.rand(p_.shape, generator=g) * 0.5 + 0.25)


def net_variant_cases():
    """Atrous residual blocks (ResBlockA) and PSP pooling (pssr/models/_blocks.py:43-92) through the reference's own modules."""
    import pssr.models as RM
    out = {}
    rng = np.random.default_rng(11)
    for tag, cls, kw, shape in VARIANT_CASES:
        torch.manual_seed(4321)
        m = getattr(RM, cls)(**kw).eval()
        randomise_variant(m)
        x = torch.tensor(rng.integers(0, 256, shape).astype(np.float32))
        with torch.no_grad():
            y = m(x)
        out[f"{tag}_x"] = x.numpy()
        out[f"{tag}_y"] = y.numpy()
        out[f"{tag}_wsum"] = np.array([float(sum(p.double().sum() for p in m.state_dict().values() if p.is_floating_point()))])
    np.savez_compressed(os.path.join(OUT, "net_variants.npz"), **out)


SWINIR_CASES = [("swinir_small", dict(image_size=32, depths=[2, 2], num_heads=[6, 6]), (2, 1, 32, 32)),
                ("swinir_w4_s2", dict(image_size=64, depths=[3], num_heads=[4], embed_dim=64, scale=2, channels=[3, 1], window_size=4), (1, 3, 40, 24))]


def randomise_swinir(m):
    """Biases and relative position tables away from their zero / tiny initial values, so that they matter."""
    g = torch.Generator().manual_seed(2)
    with torch.no_grad():
        for n, p_ in m.named_parameters():
            if "relative_position_bias_table" in n or n.endswith("bias"):
                p_.copy_(torch.randn(p_.shape, generator=g) * 0.2)


def swinir_cases():
    """SwinIR (pssr/models/swinir.py) through the reference's own module (timm's to_2tuple / trunc_normal_ / DropPath from the shim)."""
    from pssr.models import SwinIR
    out = {}
    rng = np.random.default_rng(12)
    for tag, kw, shape in SWINIR_CASES:
        torch.manual_seed(777)
        m = SwinIR(**kw).eval()
        randomise_swinir(m)
        x = torch.tensor(rng.integers(0, 256, shape).astype(np.float32))
        with torch.no_grad():
            y = m(x)
        out[f"{tag}_x"] = x.numpy()
        out[f"{tag}_y"] = y.numpy()
        out[f"{tag}_wsum"] = np.array([float(sum(p.double().sum() for p in m.state_dict().values() if p.is_floating_point()))])
    np.savez_compressed(os.path.join(OUT, "swinir.npz"), **out)


def _pillow_imread(path):
    """Stand-in for tifffile.imread in THIS script only (tifffile is absent): multi-frame grayscale TIFF -> [frames, H, W]."""
    from PIL import Image
    im = Image.open(path)
    fr = []
    for k in range(getattr(im, "n_frames", 1)):
        im.seek(k)
        fr.append(np.asarray(im))
    return np.stack(fr) if len(fr) > 1 else fr[0]


def _write_stack(path, arr):
    from PIL import Image
    ims = [Image.fromarray(a) for a in arr]
    ims[0].save(path, save_all=True, append_images=ims[1:])


def paired_cases():
    """the reference's own PairedImageDataset / PairedSlidingDataset read from TIFF files (pssr/data.py:268-431,
    `_transform_pair` :497-516) with seeded rotation draws, and `_Crappifier_Objective.sample` (pssr/train.py:348-386) with a
    deterministic crappifier on those pairs."""
    import random
    import tempfile
    from pssr import train as RT
    RD.tifffile.imread = _pillow_imread
    rng = np.random.default_rng(77)
    out = {}
    with tempfile.TemporaryDirectory() as d:
        for sub in ("ih", "il", "sh", "sl"):
            os.makedirs(os.path.join(d, sub))
        # pre-tiled pairs: 3 stacks of 6 frames, HR 56x72 (cropped to 56, reflect-padded to 64), LR 14x18 (cropped to 14, padded to 16)
        for i in range(3):
            hr = rng.integers(0, 256, (6, 56, 72)).astype(np.uint8)
            lr = rng.integers(0, 256, (6, 14, 18)).astype(np.uint8)
            _write_stack(os.path.join(d, "ih", f"im{i}.tif"), hr)
            _write_stack(os.path.join(d, "il", f"im{i}.tif"), lr)
            out[f"img_hr_{i}"], out[f"img_lr_{i}"] = hr, lr
        for tag, nf in (("a", -1), ("b", [3, 1]), ("c", 2)):
            ds = RD.PairedImageDataset(os.path.join(d, "ih"), os.path.join(d, "il"), hr_res=64, lr_scale=4, n_frames=nf, val_split=0.34,
                                       rotation=True)
            out[f"img_{tag}_len"] = np.array([len(ds)])
            out[f"img_{tag}_val"] = np.asarray(ds.val_idx)
            random.seed(5)
            for i in range(len(ds)):
                h, l = ds[i]
                out[f"img_{tag}_hr_{i}"], out[f"img_{tag}_lr_{i}"] = h.numpy(), l.numpy()
        # sheets: 2 sheets of 4 frames, HR 96x128 -> tiles 64 / overlap 32, LR 24x32 -> tiles 16 / stride 8
        for i in range(2):
            hr = rng.integers(0, 256, (4, 96, 128)).astype(np.uint8)
            lr = rng.integers(0, 256, (4, 24, 32)).astype(np.uint8)
            _write_stack(os.path.join(d, "sh", f"sh{i}.tif"), hr)
            _write_stack(os.path.join(d, "sl", f"sh{i}.tif"), lr)
            out[f"sheet_hr_{i}"], out[f"sheet_lr_{i}"] = hr, lr
        for tag, nf in (("a", [2, 1]), ("b", 1)):
            ds = RD.PairedSlidingDataset(os.path.join(d, "sh"), os.path.join(d, "sl"), hr_res=64, lr_scale=4, overlap=32, n_frames=nf,
                                         val_split=0.25, rotation=True)
            out[f"sheet_{tag}_len"] = np.array([len(ds)])
            out[f"sheet_{tag}_val"] = np.asarray(ds.val_idx)
            random.seed(6)
            for i in range(len(ds)):
                h, l = ds[i]
                out[f"sheet_{tag}_hr_{i}"], out[f"sheet_{tag}_lr_{i}"] = h.numpy(), l.numpy()
            out[f"sheet_{tag}_names"] = np.asarray([ds._get_name(i) for i in range(len(ds))])

        # the crappifier objective on image pairs whose LR half is a noisy downscale of the HR half
        class Shift(RC.Crappifier):
            """Deterministic stand-in crappifier: adds `amount` on a checkerboard (no random draws)."""
            def __init__(self, amount):
                self.amount = amount
            def crappify(self, image):
                yy, xx = np.mgrid[0:image.shape[-2], 0:image.shape[-1]]
                return image.astype(np.float64) + self.amount * ((yy + xx) % 2)

        class Pairs:
            def __init__(self, items):
                self.items = items
            def __len__(self):
                return len(self.items)
            def __getitem__(self, i):
                return self.items[i]

        items = []
        for i in range(4):
            hr = rng.integers(0, 256, (1, 64, 64)).astype(np.float32)
            lr = np.clip(rng.normal(hr[:, ::4, ::4], 9.0), 0, 255).astype(np.uint8).astype(np.float32)
            items.append((torch.as_tensor(hr), torch.as_tensor(lr)))
            out[f"obj_hr_{i}"], out[f"obj_lr_{i}"] = hr, lr
        for k, amount in enumerate((0.0, 7.0, 23.5)):
            random.seed(9)
            out[f"obj_loss_{k}"] = np.array([RT._Crappifier_Objective(Shift, Pairs(items), 4).sample([amount]), amount])
    def compact(a):          # pair tensors hold integer values 0..255 as float32: store them as uint8 (the test compares values)
        a = np.asarray(a)
        if a.dtype == np.float32 and a.size and np.array_equal(a, np.clip(np.rint(a), 0, 255)):
            return a.astype(np.uint8)
        return a
    np.savez_compressed(os.path.join(OUT, "paired.npz"), **{k: compact(v) for k, v in out.items()})


def collage_cases():
    """`_collage_preds` (pssr/predict.py:213-243) with and without normalisation -- the second `normalize_preds` call inside it takes
    the differing-resolution branch (pssr/util.py:179, skimage.transform.resize served by oracle/thirdparty.py through scipy) --
    and `normalize_preds(hr, lr)` on its own."""
    from pssr.predict import _collage_preds
    rng = np.random.default_rng(11)
    yy, xx = np.mgrid[0:64, 0:64]
    hr = np.stack([(120 + 70 * np.sin(yy / (5.0 + k)) * np.cos(xx / 7.0) + rng.normal(0, 6, (64, 64))).clip(0, 255) for k in range(2)])[:, None]
    hr = hr.astype(np.uint8)
    hat = (hr.astype(np.float64) * 0.8 + 20 + rng.normal(0, 4, hr.shape)).clip(0, 255).astype(np.uint8)
    lr = (hr[:, :, ::4, ::4].astype(np.float64) * 0.6 + 35 + rng.normal(0, 12, (2, 1, 16, 16))).clip(0, 255).astype(np.uint8)
    out = {"hr": hr, "hat": hat, "lr": lr}
    t = lambda a: torch.as_tensor(a.astype(np.float32))
    for norm in (False, True):
        im = _collage_preds(t(lr), t(hat), t(hr), norm, 5, 64, 4)
        out[f"collage_norm{int(norm)}"] = np.asarray(im)
    im = _collage_preds(t(lr), t(hat), None, False, 1, 64, 4)
    out["collage_lr_mode"] = np.asarray(im)
    a, b = RU.normalize_preds(hr, lr)
    out["norm_hr"], out["norm_lr"] = a, b
    np.savez_compressed(os.path.join(OUT, "collage.npz"), **out)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "rotation":      # added in round 2: does not touch the earlier fixtures
        gen_pair_rotation_cases()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "collage":
        collage_cases()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "paired":
        paired_cases()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "swinir":
        swinir_cases()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "variants":
        net_variant_cases()
        sys.exit(0)
    gen_pair_cases()
    gen_pair_rotation_cases()
    tiling_stitch_cases()
    normalize_cases()
    net_cases()
    paired_cases()
    collage_cases()
    net_variant_cases()
    swinir_cases()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
