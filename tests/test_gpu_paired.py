"""the paired datasets (pssr/data.py:268-431, `_transform_pair` :497-516) and the crappifier objective
(pssr/train.py:348-386) against golden vectors produced by the UNMODIFIED reference (tests/golden/gen_golden.py `paired_cases`)."""
import os
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _g():
    return np.load(os.path.join(G, "paired.npz"))


@pytest.mark.parametrize("tag,nf", [("a", -1), ("b", [3, 1]), ("c", 2)])
def test_paired_image_dataset_matches_reference_golden(tag, nf):
    from pssr2_b200.data import PairedImageDataset
    g = _g()
    hr = {f"im{i}": g[f"img_hr_{i}"] for i in range(3)}
    lr = {f"im{i}": g[f"img_lr_{i}"] for i in range(3)}
    ds = PairedImageDataset(hr, lr, hr_res=64, lr_scale=4, n_frames=nf, val_split=0.34, rotation=True)
    assert len(ds) == int(g[f"img_{tag}_len"][0]) and list(ds.val_idx) == list(g[f"img_{tag}_val"])
    assert not ds.is_lr and ds.crop_res == 64
    random.seed(5)                                  # the reference drew (getrandbits, choice) per non-validation item in this order
    for i in range(len(ds)):
        h, l = ds[i]
        assert h.dtype == torch.float32 and l.dtype == torch.float32
        assert np.array_equal(h.cpu().numpy(), g[f"img_{tag}_hr_{i}"].astype(np.float32)), (tag, i)
        assert np.array_equal(l.cpu().numpy(), g[f"img_{tag}_lr_{i}"].astype(np.float32)), (tag, i)
    with pytest.raises(IndexError):
        ds[len(ds)]
    # a batch = the same items stacked (validation items: never augmented)
    b = ds.batch(list(ds.val_idx))
    for k, i in enumerate(ds.val_idx):
        assert np.array_equal(b["hr"][k].cpu().numpy(), g[f"img_{tag}_hr_{i}"].astype(np.float32))
        assert np.array_equal(b["lr"][k].cpu().numpy(), g[f"img_{tag}_lr_{i}"].astype(np.float32))


@pytest.mark.parametrize("tag,nf", [("a", [2, 1]), ("b", 1)])
def test_paired_sliding_dataset_matches_reference_golden(tag, nf):
    from pssr2_b200.data import PairedSlidingDataset
    g = _g()
    hr = {f"sh{i}": g[f"sheet_hr_{i}"] for i in range(2)}
    lr = {f"sh{i}": g[f"sheet_lr_{i}"] for i in range(2)}
    ds = PairedSlidingDataset(hr, lr, hr_res=64, lr_scale=4, overlap=32, n_frames=nf, val_split=0.25, rotation=True)
    assert len(ds) == int(g[f"sheet_{tag}_len"][0]) and list(ds.val_idx) == list(g[f"sheet_{tag}_val"])
    assert [ds._get_name(i) for i in range(len(ds))] == list(g[f"sheet_{tag}_names"])
    random.seed(6)
    for i in range(len(ds)):
        h, l = ds[i]
        assert np.array_equal(h.cpu().numpy(), g[f"sheet_{tag}_hr_{i}"].astype(np.float32)), (tag, i)
        assert np.array_equal(l.cpu().numpy(), g[f"sheet_{tag}_lr_{i}"].astype(np.float32)), (tag, i)
    n = 0
    for h, l in ds.loader(5, train=True):
        assert h.shape[1:] == (1, 64, 64) and l.shape[2:] == (16, 16)
        n += h.shape[0]
    assert n == len(ds) - len(ds.val_idx)


def test_paired_dataset_errors():
    from pssr2_b200.data import PairedImageDataset
    a = {"x": np.zeros((1, 32, 32), np.uint8)}
    with pytest.raises(FileNotFoundError):
        PairedImageDataset(a, {"x": a["x"], "y": a["x"]}, hr_res=32, lr_scale=4)


def test_profile_hist_matches_numpy():
    from pssr2_b200 import ops
    rng = np.random.default_rng(0)
    base = rng.integers(0, 256, (3, 40, 52)).astype(np.uint8)
    for a in (rng.integers(0, 256, base.shape).astype(np.uint8), rng.normal(base, 20.0).astype(np.float32), rng.normal(base, 300.0)):
        hist, total = ops.profile_hist(torch.as_tensor(a).cuda(), torch.as_tensor(base).cuda())
        prof = a.astype(np.float32) - base.astype(np.float32)
        want, _ = np.histogram(prof.flatten(), np.arange(-256, 256))
        assert np.array_equal(hist.cpu().numpy(), want)
        assert abs(float(total.cpu()[0]) - float(prof.astype(np.float64).sum())) <= 1e-6 * max(1.0, abs(float(prof.astype(np.float64).sum())))


def test_crappifier_objective_matches_reference_golden():
    """`_Crappifier_Objective.sample` with a deterministic host crappifier: the same loss as the reference (whose float32 means differ
    from the exact sums by rounding only); a device crappifier (Poisson) runs its noise chain on the GPU and gives a finite loss
    that is smaller near the true noise level than far from it."""
    from pssr2_b200.crappifiers import AdditiveGaussian, Crappifier
    from pssr2_b200.train import _Crappifier_Objective
    g = _g()

    class Shift(Crappifier):
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

    items = [(torch.as_tensor(g[f"obj_hr_{i}"].astype(np.float32)).cuda(), torch.as_tensor(g[f"obj_lr_{i}"].astype(np.float32)).cuda()) for i in range(4)]
    for k in range(3):
        want, amount = (float(v) for v in g[f"obj_loss_{k}"])
        random.seed(9)
        got = _Crappifier_Objective(Shift, Pairs(items), 4).sample([amount])
        assert abs(got - want) <= 1e-5 * max(1.0, abs(want)), (k, got, want)
    # pairs whose LR half is the Pillow downscale of the HR half + N(0, 9) noise: the loss is smaller near sigma = 9 than far from it
    from pssr2_b200 import ops
    rng = np.random.default_rng(1)
    noisy = []
    for i in range(4):
        yy, xx = np.mgrid[0:128, 0:128]
        hr = (120 + 60 * np.sin(yy / (9.0 + i)) * np.cos(xx / 13.0)).astype(np.uint8)[None]
        ds_hr = ops.resize_bilinear(torch.as_tensor(hr).cuda(), 4).cpu().numpy().astype(np.float64)
        lr = np.clip(np.rint(ds_hr + rng.normal(0, 9.0, ds_hr.shape)), 0, 255).astype(np.float32)
        noisy.append((torch.as_tensor(hr.astype(np.float32)).cuda(), torch.as_tensor(lr).cuda()))
    obj = _Crappifier_Objective(lambda intensity: AdditiveGaussian(intensity, spread=0), Pairs(noisy), 4)
    near, far, none = obj.sample([9.0]), obj.sample([40.0]), obj.sample([0.5])
    assert np.isfinite(near) and np.isfinite(far) and near < far and near < none


def test_normalize_preds_differing_resolutions_and_collage_match_reference_golden(tmp_path):
    """pssr/util.py:179 (hr_hat at a lower resolution than hr) and `_collage_preds` / `predict_collage` (pssr/predict.py:85-142,
    :213-243).  The kernel evaluates the enlargement and the statistics exactly where the reference rounds float32 intermediates,
    so single pixels may sit one grey level apart (the bar of the equal-resolution normalisation: <= 1 LSB, >= 99.5 % identical)."""
    from pssr2_b200.predict import _collage_preds, predict_collage
    from pssr2_b200.util import normalize_preds
    from oracle import pipeline as OP
    g = np.load(os.path.join(G, "collage.npz"))
    hr, hat, lr = g["hr"], g["hat"], g["lr"]

    def close(a, b):
        d = np.abs(a.astype(np.int32) - b.astype(np.int32))
        return d.max() <= 1 and (d == 0).mean() >= 0.995

    a, b = normalize_preds(hr, lr)
    assert a.shape == hr.shape and b.shape == lr.shape and a.dtype == np.uint8
    assert close(a, g["norm_hr"]) and close(b, g["norm_lr"])
    c, d = OP.normalize_preds(hr, lr)                                    # the oracle restatement agrees with the golden bit for bit
    assert np.array_equal(c, g["norm_hr"]) and np.array_equal(d, g["norm_lr"])
    with pytest.raises(NotImplementedError):
        normalize_preds(lr, hr)
    cu = lambda x: torch.as_tensor(x).cuda()
    for norm in (False, True):
        im = np.asarray(_collage_preds(cu(lr), cu(hat), cu(hr), norm, 5, 64, 4))
        want = g[f"collage_norm{int(norm)}"]
        assert im.shape == want.shape == (128, 192)
        assert np.array_equal(im, want) if not norm else close(im, want)
    im = np.asarray(_collage_preds(cu(lr), cu(hat), None, False, 1, 64, 4))
    assert np.array_equal(im, g["collage_lr_mode"])

    # predict_collage end to end: file name, layout (input | prediction | ground truth), evaluation order
    from pssr2_b200.data import ImageDataset
    from pssr2_b200.models import ResUNet
    torch.manual_seed(3)
    model = ResUNet(hidden=[64, 128, 256], depth=1).eval()
    rng = np.random.default_rng(5)
    imgs = [rng.integers(0, 256, (1, 64, 64)).astype(np.uint8) for _ in range(5)]
    ds = ImageDataset(imgs, hr_res=64, lr_scale=4, n_frames=1, val_split=0.6, crappifier=None)
    seen = []
    predict_collage(model, ds, device="cuda", norm=False, n_images=2, prefix="t", out_dir=str(tmp_path), callbacks=[lambda: seen.append(1)])
    from PIL import Image
    im = np.asarray(Image.open(os.path.join(str(tmp_path), "t_collage_2.png")))
    assert im.shape == (128, 192)
    order = list(ds.val_idx)
    np.random.seed(0)
    np.random.shuffle(order)
    assert np.array_equal(im[:64, 128:], imgs[order[0]][0])               # third column: the ground truth, untouched without norm
    lr0 = ds.__getitem__(order[0], pp=True)[1].cpu().numpy()[0].clip(0, 255).astype(np.uint8)
    assert np.array_equal(im[:64, :64], np.repeat(np.repeat(lr0, 4, 0), 4, 1))
    assert len(seen) >= 1
    with pytest.raises(ValueError):
        predict_collage(model, ImageDataset(imgs, hr_res=64, lr_scale=-1, n_frames=1, val_split=1, crappifier=None), device="cuda", norm=True,
                        out_dir=str(tmp_path))
