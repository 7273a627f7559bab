"""Device-resident sheet pipeline (pssr2_b200.predict.predict_sheets / TilePreds), file-backed datasets with streamed residency,
the asynchronous TIFF writer, and the per-item rules of `batch()` (frame windows, spread)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _model():
    from pssr2_b200.models import ResUNet
    torch.manual_seed(3)
    return ResUNet(hidden=[64, 128, 256], depth=1).eval()


def _sheets(n=3, frames=1, shape=(448, 640), dtype=np.uint16):
    out = {}
    for i in range(n):
        rng = np.random.default_rng(i)
        yy, xx = np.mgrid[0:shape[0], 0:shape[1]]
        base = 90 + 60 * np.sin(yy / (17.0 + i)) * np.cos(xx / 31.0)
        out[f"sheet{i}"] = rng.poisson(np.broadcast_to(base, (frames,) + shape)).clip(0, 255).astype(dtype)
    return out


@pytest.mark.parametrize("frames,n_frames,margin", [(1, -1, 8), (4, 2, 0), (1, -1, 40)])
def test_predict_sheets_equals_predict_then_reassemble(frames, n_frames, margin):
    from pssr2_b200.data import SlidingDataset
    from pssr2_b200.predict import TilePreds, predict_images, predict_sheets
    from pssr2_b200.util import reassemble_sheets
    model = _model()
    sheets = _sheets(3, frames)
    kw = dict(hr_res=256, lr_scale=4, overlap=64, val_split=1, crappifier=None, n_frames=n_frames)
    if frames > 1:      # multi-frame input: the model takes n_frames channels
        from pssr2_b200.models import ResUNet
        torch.manual_seed(3)
        model = ResUNet(channels=[2, 1], hidden=[64, 128, 256], depth=1).eval()
    ds = SlidingDataset(sheets, **kw)
    preds = predict_images(model, ds, device="cuda", batch_size=5, out_dir=None)
    want = reassemble_sheets(preds, ds, lr_scale=1, overlap=64, margin=margin, out_dir=None)
    got = predict_sheets(model, SlidingDataset(sheets, **kw), device="cuda", batch_size=5, overlap=64, margin=margin)
    assert len(got) == len(want) == 3
    for a, b in zip(got, want):
        assert a.dtype == np.uint8 and a.shape == b.shape and np.array_equal(a, b)
    # two-call flow with the tiles kept in HBM
    dp = predict_images(model, SlidingDataset(sheets, **kw), device="cuda", batch_size=5, out_dir=None, keep_on_device=True)
    assert isinstance(dp, TilePreds) and sorted(dp) == sorted(preds) and len(dp) == len(preds)
    name = next(iter(preds))
    assert np.array_equal(dp[name], preds[name]) and dict(dp.items())[name].shape == preds[name].shape
    again = reassemble_sheets(dp, ds, lr_scale=1, overlap=64, margin=margin, out_dir=None)
    assert all(np.array_equal(a, b) for a, b in zip(again, want))
    with pytest.raises(ValueError):
        predict_images(model, ds, device="cuda", out_dir="x", keep_on_device=True)


def test_file_backed_dataset_streams_like_memory(tmp_path):
    """A directory of TIFF sheets (decoded by the library's reader into pinned memory on the reader thread) gives the same
    predictions as in-memory arrays -- also when only one sheet may be resident at a time -- and `out_dir` writes the same
    images through the asynchronous TIFF writer."""
    from pssr2_b200 import io
    from pssr2_b200.data import SlidingDataset
    from pssr2_b200.predict import predict_images, predict_sheets
    model = _model()
    sheets = _sheets(4)
    src = tmp_path / "sheets"          # (the datasets glob their directory recursively, like the reference: outputs live elsewhere)
    src.mkdir()
    for k, v in sheets.items():
        io.write_tiff(src / f"{k}.tif", v)
    kw = dict(hr_res=256, lr_scale=4, overlap=64, val_split=1, crappifier=None)
    want = predict_images(model, SlidingDataset(sheets, **kw), device="cuda", batch_size=4, out_dir=None)
    ds = SlidingDataset(str(src), extension="tif", **kw)
    ds.max_resident_bytes = sheets["sheet0"].nbytes          # room for one sheet: the others stream through
    got = predict_images(model, ds, device="cuda", batch_size=4, out_dir=None)
    assert sorted(got) == sorted(want) and all(np.array_equal(got[k], want[k]) for k in want)
    assert sum(s is not None for s in ds._sheets) <= 2
    # no explicit budget: a dataset larger than half of the free HBM gets one by itself (here: "free" memory of 2.5 sheets)
    real = torch.cuda.mem_get_info
    try:
        torch.cuda.mem_get_info = lambda *a, **k: (int(2.5 * sheets["sheet0"].nbytes), real()[1])
        ds_auto = SlidingDataset(sheets, **kw)
    finally:
        torch.cuda.mem_get_info = real
    assert ds_auto.max_resident_bytes == int(0.5 * int(2.5 * sheets["sheet0"].nbytes)) and sum(s is not None for s in ds_auto._sheets) == 1
    got = predict_images(model, ds_auto, device="cuda", batch_size=4, out_dir=None)
    assert all(np.array_equal(got[k], want[k]) for k in want) and sum(s is not None for s in ds_auto._sheets) <= 2
    out = tmp_path / "preds"
    assert predict_images(model, SlidingDataset(str(src), extension="tif", **kw), device="cuda", batch_size=3, out_dir=str(out), prefix="p") is None
    files = sorted(os.listdir(out))
    assert len(files) == len(want) and all(f.startswith("p_") for f in files)
    for k in want:
        assert np.array_equal(io.read_tiff(out / f"p_{k}.tif"), want[k])
    sd = tmp_path / "stitched"
    predict_sheets(model, SlidingDataset(str(src), extension="tif", **kw), device="cuda", batch_size=4, margin=8, out_dir=str(sd))
    mem = predict_sheets(model, SlidingDataset(sheets, **kw), device="cuda", batch_size=4, margin=8)
    for i in range(4):
        assert np.array_equal(io.read_tiff(sd / f"sheet{i}.tif"), mem[i])


def test_batch_rules_frame_windows_and_spread():
    from pssr2_b200 import ops
    from pssr2_b200.crappifiers import AdditiveGaussian, MultiCrappifier, Poisson
    from pssr2_b200.data import ImageDataset
    a = np.random.default_rng(0).integers(0, 255, (3, 64, 64)).astype(np.uint8)
    b = np.random.default_rng(1).integers(0, 255, (5, 64, 64)).astype(np.uint8)
    ds = ImageDataset([a, b], hr_res=64, lr_scale=4, n_frames=-1, val_split=1, crappifier=None)
    with pytest.raises(ValueError):
        ds.batch([0, 1])                       # 3-frame and 5-frame stacks cannot form one batch tensor
    assert ds.batch([0])["lr"].shape == (1, 3, 16, 16) and ds.batch([1])["lr"].shape == (1, 5, 16, 16)
    table = ops.TileTable([torch.as_tensor(a).cuda()], [0], [1], [0], [0], [64], [64])
    with pytest.raises(ValueError):
        ops.crappify(table, 64, 4, None, frames=3)            # frames [1, 4) of a 3-frame sheet: rejected on the host
    # spread > 0: one intensity draw PER ITEM (pssr/crappifiers.py:63,85), not per batch
    crap = MultiCrappifier(Poisson(spread=0.1), AdditiveGaussian(spread=2.0))
    ds2 = ImageDataset([a[:1], a[1:2], a[2:3]], hr_res=64, lr_scale=4, n_frames=1, val_split=1, crappifier=crap)
    calls = []
    real = np.random.normal
    np.random.normal = lambda *aa, **kk: (calls.append(1), real(*aa, **kk))[1]
    try:
        out = ds2.batch([0, 1, 2])
    finally:
        np.random.normal = real
    assert out["lr"].shape == (3, 1, 16, 16) and len(calls) == 6       # 3 items x 2 stages


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_second_device_while_first_is_current():
    """Tensors on cuda:1 with device 0 current: every library call must run in device 1's context."""
    from pssr2_b200.data import SlidingDataset
    from pssr2_b200.predict import predict_images, test_metrics as run_metrics
    assert torch.cuda.current_device() == 0
    model = _model()
    sheets = _sheets(1)
    kw = dict(hr_res=256, lr_scale=4, overlap=64, val_split=1, crappifier=None)
    want = predict_images(model, SlidingDataset(sheets, device="cuda:0", **kw), device="cuda:0", batch_size=4, out_dir=None)
    got = predict_images(model, SlidingDataset(sheets, device="cuda:1", **kw), device="cuda:1", batch_size=4, out_dir=None)
    assert torch.cuda.current_device() == 0
    assert all(np.array_equal(got[k], want[k]) for k in want)
    m = run_metrics(model, SlidingDataset(sheets, device="cuda:1", **kw), device="cuda:1", norm=True)
    assert all(np.isfinite(v) for v in m.values())


def test_training_augmentation_matches_oracle():
    """SURVEY 8f-2, the data side of train_paired: rot90 / flip of the padded tile before the downscale (pssr/data.py:476-480),
    bit-exact against the oracle for every transform code, HR and LR, uint8 / uint16, with a reflect-padded tile; `loader`
    walks the training items only and augments them, validation items never rotate."""
    import random
    from oracle import pipeline as OP
    from pssr2_b200 import ops
    from pssr2_b200.data import ImageDataset
    rng = np.random.default_rng(3)
    for dt, shape, hr_res, scale in ((np.uint8, (2, 64, 64), 64, 4), (np.uint16, (1, 50, 58), 64, 2)):
        img = rng.integers(0, 255, shape).astype(dt)
        for rot in (False, True):
            for axes in (1, 2, (1, 2)):
                ax = (axes,) if isinstance(axes, int) else axes
                code = (1 if rot else 0) | (2 if 1 in ax else 0) | (4 if 2 in ax else 0)
                F, h, w = img.shape
                size = min(h, w, hr_res)
                y0, x0 = ((h - size) // 2, (w - size) // 2) if [h, w] != [hr_res] * 2 else (0, 0)
                t = torch.as_tensor(img.view(np.int16) if dt == np.uint16 else img).cuda()
                table = ops.TileTable([t], [0], [0], [y0], [x0], [size], [size], tile_xf=[code])
                lr, hr, hr8 = ops.crappify(table, hr_res, scale, None, frames=F, want_hr_f32=True, want_hr_u8=True)
                whr, wlr = OP.gen_pair(img, hr_res, scale, None, rotation=[rot, axes])
                assert np.array_equal(hr.cpu().numpy()[0], whr), (dt, rot, axes)
                assert np.array_equal(lr.cpu().numpy()[0], wlr), (dt, rot, axes)
                assert np.array_equal(hr8.cpu().numpy()[0, 0], np.clip(whr[F // 2], 0, 255).astype(np.uint8))
    imgs = [rng.integers(0, 255, (1, 64, 64)).astype(np.uint8) for _ in range(10)]
    ds = ImageDataset(imgs, hr_res=64, lr_scale=4, n_frames=1, val_split=0.2, crappifier=None, rotation=True)
    assert len(ds.val_idx) == 2
    plain = {i: ds.__getitem__(i, pp=True) for i in range(10)}
    for i in ds.val_idx:                                   # validation items: never augmented
        assert torch.equal(ds[i][0], plain[i][0])
    random.seed(1)
    seen, changed = 0, 0
    for hr, lr in ds.loader(3, train=True):
        assert hr.shape[1:] == (1, 64, 64) and lr.shape[1:] == (1, 16, 16)
        seen += hr.shape[0]
    assert seen == 8
    random.seed(2)
    for i in range(10):
        if i not in ds.val_idx:
            changed += int(not torch.equal(ds[i][0], plain[i][0]))
    assert changed >= 5                                    # 7 of the 8 transform codes move pixels
    assert sum(b[0].shape[0] for b in ds.loader(4, train=False)) == 2
