"""GPU end-to-end parity of the reference-facing entry points (predict_images / test_metrics /
reassemble_sheets over the device datasets) against the CPU oracle chained the way the reference
chains it (pssr/predict.py:11-83, :144-211).  Noise-free (crappifier=None) so the comparison is exact
up to the network's 16-bit operand rounding: predictions may differ by +-1 on a small fraction of
pixels (truncation boundaries); metrics must agree within 1e-3 (north_star tolerance)."""
import numpy as np
import pytest
import torch

from oracle import pipeline as OP
from oracle.models import resunet_forward

pytestmark = pytest.mark.gpu


def _model():
    from pssr2_b200.models import ResUNet
    torch.manual_seed(3)
    m = ResUNet(hidden=[64, 128, 256], depth=1).eval()
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    return m, sd


def _sheet(dtype=np.uint16, shape=(1, 448, 448), seed=0):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:shape[1], 0:shape[2]]
    base = 90 + 60 * np.sin(yy / 23.0) * np.cos(xx / 31.0)
    return rng.poisson(np.broadcast_to(base, shape)).clip(0, 255).astype(dtype)


def test_predict_images_and_stitch_vs_oracle():
    from pssr2_b200.data import SlidingDataset
    from pssr2_b200.predict import predict_images
    from pssr2_b200.util import reassemble_sheets
    model, sd = _model()
    sheet = _sheet()
    ds = SlidingDataset({"s0": sheet}, hr_res=256, lr_scale=4, overlap=64, val_split=1, crappifier=None)
    assert len(ds) == 4 and ds.val_idx == [0, 1, 2, 3]
    preds = predict_images(model, ds, device="cuda", batch_size=3, out_dir=None)
    assert sorted(preds) == ["s0_0_0", "s0_1_0", "s0_2_0", "s0_3_0"]
    want_tiles = []
    for i in range(4):
        hr = OP.sliding_window(sheet, 256, 192, None, 1, i)
        _, lr = OP.gen_pair(hr, 256, 4, None)
        out = resunet_forward(sd, torch.as_tensor(lr)[None])
        w = OP.pred_array(out.numpy())[0]
        want_tiles.append(w[0])
        g = preds[f"s0_{i}_0"]
        assert g.shape == (1, 256, 256) and g.dtype == np.uint8
        d = np.abs(g.astype(int) - w.astype(int))
        assert d.max() <= 1 and (d != 0).mean() < 0.02, (d.max(), (d != 0).mean())
    sheets = reassemble_sheets(preds, ds, lr_scale=1, overlap=64, margin=8, out_dir=None)
    got_tiles = np.stack([preds[f"s0_{i}_0"][0] for i in range(4)])
    assert np.array_equal(sheets[0], OP.stitch_sheets(got_tiles, 2, 2, 64, 8))     # stitch itself is bit-exact


def test_test_metrics_vs_oracle():
    from pssr2_b200.data import ImageDataset
    from pssr2_b200.predict import test_metrics as run_metrics
    model, sd = _model()
    imgs = _sheet(np.uint8, (3, 256, 256), seed=4)
    ds = ImageDataset([imgs], hr_res=256, lr_scale=4, n_frames=1, val_split=1, crappifier=None)
    assert len(ds) == 3
    for norm in (True, False):
        got = run_metrics(model, ds, device="cuda", norm=norm, avg=False, item0_quirk=False, batch_size=2)
        for i in range(3):
            hr, lr = OP.gen_pair(imgs[i:i + 1], 256, 4, None)
            out = resunet_forward(sd, torch.as_tensor(lr)[None])
            a, b = OP.pred_array(hr[None]), OP.pred_array(out.numpy())
            if norm:
                a, b = OP.normalize_preds(a, b)
            mse, pixel, psnr, ssim = OP.image_metrics(a[0], b[0])
            print(f"norm={norm} img{i}: mse {got['mse'][i]:.6g}/{mse:.6g} psnr {got['psnr'][i]:.5f}/{psnr:.5f} ssim {got['ssim'][i]:.6f}/{ssim:.6f}")
            assert abs(got["mse"][i] - mse) <= 1e-3 * mse + 1e-9
            assert abs(got["pixel"][i] - pixel) <= 1e-3 * pixel + 1e-6
            assert abs(got["psnr"][i] - psnr) <= 1e-3
            assert abs(got["ssim"][i] - ssim) <= 1e-3
    # reference quirk: every iteration scores dataset[0] (pssr/predict.py:180)
    q = run_metrics(model, ds, device="cuda", norm=False, avg=False)
    assert len(q["mse"]) == 3 and max(q["mse"]) - min(q["mse"]) < 1e-12
    avg = run_metrics(model, ds, device="cuda", norm=False)
    assert set(avg) == {"mse", "pixel", "psnr", "ssim"} and all(isinstance(v, float) for v in avg.values())


def test_crappifier_operator_interface():
    """reference tests/test_crappifiers.py: shape in == shape out, for every class and kwargs grid."""
    from pssr2_b200.crappifiers import AdditiveGaussian, MultiCrappifier, Poisson, SaltPepper
    rng = np.random.default_rng(0)
    img = rng.random((2, 1, 128, 128)) * 255
    for cls in (AdditiveGaussian, Poisson, SaltPepper):
        for kw in ({}, dict(intensity=2), dict(intensity=0.5), dict(gain=10), dict(gain=-10), dict(spread=0.5)):
            out = cls(**kw)(img)
            assert out.shape == img.shape and np.isfinite(out).all()
    out = MultiCrappifier(AdditiveGaussian(), Poisson(), SaltPepper())(img)
    assert out.shape == img.shape and out.min() >= 0 and out.max() <= 255
    np.random.seed(1)
    a = Poisson()(img)
    np.random.seed(1)
    assert np.array_equal(a, Poisson()(img)), "np.random.seed must make a run reproducible"
    assert abs(a.mean() - img.mean()) < 0.5


def test_dataset_shapes_like_reference_tests():
    """reference tests/test_data.py: item shapes for single-frame, multi-frame, LR mode, crop/pad."""
    from pssr2_b200.data import ImageDataset, SlidingDataset
    rng = np.random.default_rng(2)
    ims = [rng.integers(0, 256, (1, 512, 512)).astype(np.uint8) for _ in range(5)]
    ds = ImageDataset(ims)
    assert len(ds) == 5 and str(ds)
    hr, lr = ds[0]
    assert tuple(hr.shape) == (1, 512, 512) and tuple(lr.shape) == (1, 128, 128)
    ds = ImageDataset([rng.integers(0, 256, (10, 512, 512)).astype(np.uint8) for _ in range(2)], n_frames=2)
    assert len(ds) == 10
    hr, lr = ds[3]
    assert tuple(hr.shape) == (2, 512, 512) and tuple(lr.shape) == (2, 128, 128)
    ds = ImageDataset([rng.integers(0, 256, (1, 128, 128)).astype(np.uint8) for _ in range(3)], val_split=1)
    assert ds.is_lr and tuple(ds[0].shape) == (1, 128, 128)
    ds = ImageDataset([rng.integers(0, 256, (1, 500, 500)).astype(np.uint8) for _ in range(3)])
    assert ds.crop_res == 500
    hr, lr = ds[0]
    assert tuple(hr.shape) == (1, 512, 512) and tuple(lr.shape) == (1, 128, 128)
    sheets = [rng.integers(0, 256, (1, 1024, 1024)).astype(np.uint8) for _ in range(2)]
    ds = SlidingDataset(sheets, extension="tif", overlap=None, preload=False)
    assert len(ds) == 2 * 4
    hr, lr = ds[5]
    assert tuple(hr.shape) == (1, 512, 512) and tuple(lr.shape) == (1, 128, 128)
    ds = SlidingDataset([s[:, :256, :256] for s in sheets], hr_res=128, lr_scale=-1, extension="tif", overlap=None, val_split=1)
    assert ds.is_lr and len(ds) == 8 and tuple(ds[0].shape) == (1, 128, 128)


def test_error_behaviour_callbacks_and_file_output(tmp_path):
    """Reference error contract (pssr/predict.py:40, pssr/util.py:76-77), the callback protocol (pssr/util.py:228-231,
    predict.py:75-79: called after every image, with locals() iff it takes one argument; raising aborts the run), the TIFF
    output path (predict.py:66-73) and the no-fallback rule."""
    from PIL import Image
    from pssr2_b200.data import ImageDataset, SlidingDataset
    from pssr2_b200.predict import predict_images
    from pssr2_b200.util import reassemble_sheets
    model, _ = _model()
    rng = np.random.default_rng(9)
    lr_only = ImageDataset([rng.integers(0, 256, (1, 64, 64)).astype(np.uint8) for _ in range(3)], hr_res=256, lr_scale=4, val_split=1)
    assert lr_only.is_lr
    with pytest.raises(ValueError):
        predict_images(model, lr_only, device="cuda", norm=True, out_dir=None)
    with pytest.raises(RuntimeError):
        predict_images(model, lr_only, device="cpu", out_dir=None)
    with pytest.raises(ValueError):
        reassemble_sheets({}, {"s": (1, 64, 64)}, lr_scale=4, overlap=8, margin=9, out_dir=None)
    # LR-mode prediction: 64^2 -> 256^2, crop = crop_res * scale (predict.py:66)
    seen, seen_locals = [], []
    preds = predict_images(model, lr_only, device="cuda", batch_size=2, out_dir=None,
                           callbacks=[lambda: seen.append(1), lambda loc: seen_locals.append(sorted(loc)[:1])])
    assert len(preds) == 3 and all(v.shape == (1, 256, 256) and v.dtype == np.uint8 for v in preds.values())
    assert len(seen) == 3 and len(seen_locals) == 3

    class Abort(Exception):
        pass

    def stop():
        raise Abort()
    with pytest.raises(Abort):
        predict_images(model, lr_only, device="cuda", out_dir=None, callbacks=[stop])
    # file output: {out_dir}/{prefix_}{name}.tif, identical pixels to the dict path
    sheet = _sheet(np.uint8, (1, 320, 320), seed=5)
    ds = SlidingDataset({"a": sheet}, hr_res=256, lr_scale=4, overlap=192, val_split=1, crappifier=None)
    want = predict_images(model, ds, device="cuda", batch_size=3, out_dir=None)
    predict_images(model, ds, device="cuda", batch_size=3, out_dir=str(tmp_path / "p"), prefix="run")
    for name, arr in want.items():
        got = np.asarray(Image.open(tmp_path / "p" / f"run_{name}.tif"))
        assert np.array_equal(got, arr[0])
    # an empty validation split predicts nothing
    ds.val_idx = []
    assert predict_images(model, ds, device="cuda", out_dir=None) == {}


def test_heterogeneous_image_sizes():
    """Images / sheets of different sizes in one dataset (the reference crops / pads / tiles each image on its own,
    pssr/data.py:100-120, :236-256, :536-551): per-sheet dimensions travel to the gather kernel."""
    from pssr2_b200.data import ImageDataset, SlidingDataset
    from pssr2_b200.predict import predict_images
    rng = np.random.default_rng(21)
    # (100, 90): the reflect pad (166) is wider than the image, np.pad repeats the reflection
    ims = [rng.integers(0, 256, (1, h, w)).astype(np.uint8) for h, w in ((256, 256), (200, 240), (300, 280), (131, 256), (100, 90))]
    ds = ImageDataset(ims, hr_res=256, lr_scale=4, n_frames=1, val_split=1, crappifier=None)
    assert len(ds) == 5 and ds.crop_res == 256
    for i, im in enumerate(ims):
        hr, lr = ds[i]
        want_hr, want_lr = OP.gen_pair(im, 256, 4, None)
        assert np.array_equal(hr.cpu().numpy(), want_hr) and np.array_equal(lr.cpu().numpy(), want_lr), f"image {i}"
    sheets = {"a": rng.integers(0, 256, (1, 320, 448)).astype(np.uint16), "b": rng.integers(0, 256, (1, 448, 256)).astype(np.uint16)}
    sd = SlidingDataset(sheets, hr_res=128, lr_scale=4, overlap=32, val_split=1, crappifier=None)
    counts = [OP.n_tiles(s.shape[-2:], 128, 96) for s in sheets.values()]
    assert len(sd) == sum(a * b for a, b in counts)
    idx = 0
    for name, s in sheets.items():
        ta, tb = OP.n_tiles(s.shape[-2:], 128, 96)
        for t in range(ta * tb):
            hr, lr = sd[idx]
            want_hr, want_lr = OP.gen_pair(OP.sliding_window(s, 128, 96, None, 1, t), 128, 4, None)
            assert np.array_equal(hr.cpu().numpy(), want_hr) and np.array_equal(lr.cpu().numpy(), want_lr), (name, t)
            idx += 1
    model, _ = _model()
    preds = predict_images(model, sd, device="cuda", batch_size=5, out_dir=None)
    assert len(preds) == len(sd) and all(v.shape == (1, 128, 128) for v in preds.values())
