"""GPU parity tests (through the C ABI) for kernel families 1 (crappify), 3 (stitch) and 4 (scoring)
against the CPU oracle on the same seeded inputs.  Integer / byte work must be bit-exact."""
import numpy as np
import pytest
import torch

from oracle import pipeline as OP
from oracle.pillow_resize import resize_bilinear as oracle_resize

pytestmark = pytest.mark.gpu


def _ops():
    from pssr2_b200 import ops
    return ops


def _dev(a):
    return torch.as_tensor(a).cuda()


@pytest.mark.parametrize("dtype,h,scale", [(np.uint8, 512, 4), (np.uint16, 512, 4), (np.uint8, 512, 8), (np.uint16, 2048, 8),
                                           (np.uint8, 96, 3), (np.uint16, 500, 4), (np.uint8, 128, 1)])
def test_resize_bit_exact(dtype, h, scale):
    ops = _ops()
    rng = np.random.default_rng(h + scale)
    img = rng.integers(0, 256 if dtype == np.uint8 else 65536, (2, h, h)).astype(dtype)
    want = oracle_resize(img, h // scale, h // scale)
    got = ops.resize_bilinear(_dev(img), scale).cpu().numpy()
    assert got.dtype == want.dtype and np.array_equal(got, want)


def _sheet_case(dtype, n_sheets, frames, H, W, rng, maxv=256):
    return [rng.integers(0, maxv, (frames, H, W)).astype(dtype) for _ in range(n_sheets)]


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16])
@pytest.mark.parametrize("hr_res,scale,stride", [(512, 4, 384), (256, 8, 256), (128, 4, 96)])
def test_crappify_injected_bit_exact(dtype, hr_res, scale, stride):
    """SlidingDataset tiling + downscale + MultiCrappifier(Poisson, AdditiveGaussian, SaltPepper) with the
    reference's draws injected: float32 LR must equal the oracle bit for bit."""
    ops = _ops()
    rng = np.random.default_rng(7)
    frames = 2
    sheets = _sheet_case(dtype, 2, frames, 1024 if hr_res > 128 else 300, 1100 if hr_res > 128 else 417, rng)
    tx, ty = OP.n_tiles(sheets[0].shape[-2:], hr_res, stride)
    tiles = [(s, t) for s in range(len(sheets)) for t in range(tx * ty)]
    lr_res = hr_res // scale
    n = len(tiles)
    ys = rng.poisson(100, (n, frames, lr_res, lr_res)).astype(np.int64)
    gs = rng.normal(0.5, 13, (n, frames, lr_res, lr_res))
    fl = rng.random((n, frames, lr_res, lr_res)) <= 0.05
    sa = rng.random((n, frames, lr_res, lr_res)) <= 0.5
    want_lr, want_hr = [], []
    t_sheet, t_y, t_x = [], [], []
    for i, (s, t) in enumerate(tiles):
        hr = OP.sliding_window(sheets[s], hr_res, stride, frames, 1, t)
        t_sheet.append(s)
        t_y.append(t // ty * stride)
        t_x.append(t % ty * stride)
        stages = [("poisson", ys[i], 0.7, 1.5), ("gaussian", gs[i]), ("saltpepper", fl[i], sa[i], -2.0)]
        h, l = OP.gen_pair(hr, hr_res, scale, stages)
        want_lr.append(l)
        want_hr.append(h)
    table = ops.TileTable([_dev(s.view(np.int16) if dtype == np.uint16 else s) for s in sheets], t_sheet, [0] * n, t_y, t_x,
                          [hr_res] * n, [hr_res] * n)
    spec = [ops.NoiseSpec(1, 0.7, 1.5, True, _dev(ys)), ops.NoiseSpec(2, 13, 0.5, True, _dev(gs)),
            ops.NoiseSpec(3, 0.05, -2.0, True, _dev((fl.astype(np.uint8) | (sa.astype(np.uint8) << 1))))]
    lr, hr, hr8 = ops.crappify(table, hr_res, scale, spec, frames=frames, clip_between=True, want_hr_f32=True, want_hr_u8=True)
    assert np.array_equal(lr.cpu().numpy(), np.stack(want_lr)), "LR tiles differ from the oracle"
    assert np.array_equal(hr.cpu().numpy(), np.stack(want_hr)), "HR tiles differ from the oracle"
    assert np.array_equal(hr8.cpu().numpy(), OP.pred_array(np.stack(want_hr))), "uint8 HR differs"


def test_crappify_pad_no_noise_and_frames():
    """ImageDataset path: centre crop + reflect pad (500 -> 512), crappifier=None, n_frames=[3,1]."""
    ops = _ops()
    rng = np.random.default_rng(11)
    img = rng.integers(0, 256, (5, 520, 500)).astype(np.uint8)
    hr_res, scale = 512, 4
    want_hr, want_lr = OP.gen_pair(img.copy(), hr_res, scale, None, n_frames=[3, 1])
    size = min(520, 500, hr_res)
    table = ops.TileTable([_dev(img)], [0], [0], [(520 - size) // 2], [(500 - size) // 2], [size], [size])
    lr, hr, _ = ops.crappify(table, hr_res, scale, None, frames=5, lr_frame0=1, lr_frames=3, hr_frame0=2, hr_frames=1, want_hr_f32=True)
    assert np.array_equal(lr.cpu().numpy()[0], want_lr)
    assert np.array_equal(hr.cpu().numpy()[0], want_hr)


def test_crappify_philox_statistics():
    """Free-running mode cannot match MT19937; check moments of Poisson / Gaussian / salt&pepper draws."""
    ops = _ops()
    hr_res, scale = 512, 4
    for level in (3.0, 40.0, 200.0):
        img = np.full((1, hr_res, hr_res), int(level), np.uint8)
        table = ops.TileTable([_dev(img)] , [0] * 8, [0] * 8, [0] * 8, [0] * 8, [hr_res] * 8, [hr_res] * 8)
        lr, _, _ = ops.crappify(table, hr_res, scale, [ops.NoiseSpec(1, 1, 0)], seed=5)
        v = lr.double().cpu().numpy()
        n = v.size
        assert abs(v.mean() - int(level)) < 5 * np.sqrt(level / n) + 0.02, (level, v.mean())
        assert abs(v.var() - int(level)) < 0.05 * level + 0.1, (level, v.var())
        assert np.array_equal(v, np.round(v))
        lr2, _, _ = ops.crappify(table, hr_res, scale, [ops.NoiseSpec(1, 1, 0)], seed=5)
        assert torch.equal(lr, lr2), "Philox mode must be deterministic for a fixed seed"
        lr3, _, _ = ops.crappify(table.slice(4, 4), hr_res, scale, [ops.NoiseSpec(1, 1, 0)], seed=5, tile_index0=4)
        assert torch.equal(lr[4:], lr3), "noise must depend on the global tile index only (sharding invariance)"
    img = np.full((1, hr_res, hr_res), 100, np.uint8)
    table = ops.TileTable([_dev(img)], [0] * 8, [0] * 8, [0] * 8, [0] * 8, [hr_res] * 8, [hr_res] * 8)
    lr, _, _ = ops.crappify(table, hr_res, scale, [ops.NoiseSpec(2, 13, 2)], seed=9)
    v = lr.double().cpu().numpy()
    assert abs(v.mean() - 102) < 0.2 and abs(v.std() - 13) < 0.3, (v.mean(), v.std())
    lr, _, _ = ops.crappify(table, hr_res, scale, [ops.NoiseSpec(3, 0.05, 0)], seed=9)
    v = lr.cpu().numpy()
    assert abs((v == 255).mean() - 0.025) < 0.004 and abs((v == 0).mean() - 0.025) < 0.004


@pytest.mark.parametrize("n_rows,n_cols,T,ov,margin,stacks", [(3, 4, 64, 16, 0, 2), (3, 3, 64, 16, 8, 1), (2, 5, 64, 16, 12, 1),
                                                               (4, 4, 32, 20, 4, 1), (1, 1, 64, 0, 0, 3), (10, 10, 512, 128, 32, 1),
                                                               (3, 4, 64, 18, 5, 2),      # step / margin not multiples of 4: scalar path
                                                               (3, 3, 64, 40, 12, 1),     # overlap > T/2: up to 9 contributors per pixel
                                                               (4, 5, 80, 48, 16, 2),     # 16-pixel path with 3 / 6 / 9 contributors (reciprocal division)
                                                               (3, 3, 128, 32, 16, 1)])   # 16-pixel path, 1 / 2 / 4 contributors
def test_stitch_bit_exact(n_rows, n_cols, T, ov, margin, stacks):
    ops = _ops()
    rng = np.random.default_rng(n_rows * 100 + n_cols)
    tiles = rng.integers(0, 256, (stacks * n_rows * n_cols, T, T)).astype(np.uint8)
    want = OP.stitch_sheets(tiles, n_rows, n_cols, ov, margin)
    got = ops.stitch(_dev(tiles), n_rows, n_cols, ov, margin).cpu().numpy()
    assert np.array_equal(got, want)


def test_stitch_roundtrip_full_size():
    """Size-independent property at config-3 scale: tiling a sheet then stitching returns the sheet."""
    ops = _ops()
    rng = np.random.default_rng(3)
    sheet = rng.integers(0, 256, (4096, 4096)).astype(np.uint8)
    T, ov = 512, 128
    tx, ty = OP.n_tiles(sheet.shape, T, T - ov)
    tiles = np.stack([OP.sliding_window(sheet[None], T, T - ov, None, 1, t)[0] for t in range(tx * ty)])
    got = ops.stitch(_dev(tiles), tx, ty, ov, 32).cpu().numpy()[0]
    assert np.array_equal(got, sheet[:got.shape[0], :got.shape[1]])
    with pytest.raises(ValueError):
        ops.stitch(_dev(tiles), tx, ty, ov, ov + 1)


@pytest.mark.parametrize("h,w", [(512, 512), (100, 77), (7, 7), (64, 200)])
def test_metrics_match_oracle(h, w):
    ops = _ops()
    rng = np.random.default_rng(h * w)
    a = rng.integers(0, 256, (3, h, w)).astype(np.uint8)
    b = np.clip(a.astype(int) + rng.integers(-20, 21, a.shape), 0, 255).astype(np.uint8)
    b[2] = a[2]  # identical pair: mse 0, psnr inf, ssim 1
    sq, ss = ops.metric_sums(_dev(a), _dev(b))
    sq, ss = sq.cpu().numpy(), ss.cpu().numpy()
    for i in range(3):
        assert sq[i] == int(((a[i].astype(np.int64) - b[i].astype(np.int64)) ** 2).sum())
        mse, pixel, psnr, ssim = OP.image_metrics(a[i][None], b[i][None])
        # window sums are exact integers; the quotient is evaluated in float (2e-7 per window, random sign) -- tolerance of the path: 1e-3
        assert abs(ss[i] / ((h - 6) * (w - 6)) - ssim) < 1e-6, (ss[i] / ((h - 6) * (w - 6)), ssim)
        got_mse = sq[i] / (h * w) / 255.0 ** 2
        assert abs(got_mse - mse) < 1e-12


def test_normalize_preds_close_to_oracle():
    """Floating-point stage: the reference works in float32 with NumPy's pairwise sums; the kernel takes
    the same statistics exactly from histograms, so a handful of pixels may land on the other side of a
    truncation boundary.  Tolerance: |diff| <= 1 everywhere and >= 99.5 % of pixels identical."""
    ops = _ops()
    rng = np.random.default_rng(5)
    base = rng.poisson(90, (4, 256, 256)).clip(0, 255)
    hr = base.astype(np.uint8)
    hat = np.clip(base * 0.8 + 20 + rng.normal(0, 6, base.shape), 0, 255).astype(np.uint8)
    wa, wb = OP.normalize_preds(hr, hat)
    ga, gb = ops.normalize_preds_u8(_dev(hr), _dev(hat))
    ga, gb = ga.cpu().numpy(), gb.cpu().numpy()
    for w, g in ((wa, ga), (wb, gb)):
        d = np.abs(w.astype(int) - g.astype(int))
        assert d.max() <= 1 and (d == 0).mean() >= 0.995, (d.max(), (d == 0).mean())
