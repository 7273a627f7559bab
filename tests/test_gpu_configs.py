"""GPU parity at the FULL sizes of BASELINE.json configs 3, 4, 5 (config 2 is the bench workload, covered by
tests/test_gpu_net.py / test_gpu_predict.py): the CUDA path against the CPU oracle where the oracle finishes in seconds,
and size-independent properties (bit-exact stitching of the produced tiles, partition invariance, metric identities) on the
whole problem.  Tolerances as in tests/test_gpu_predict.py (north_star: tiling / crappify bit-exact, network >= 50 dB,
metrics within 1e-3)."""
import numpy as np
import pytest
import torch

from oracle import pipeline as OP

pytestmark = pytest.mark.gpu


def _psnr(a, b):
    mse = float(((a.astype(np.float64) - b.astype(np.float64)) ** 2).mean())
    return 10 * np.log10(255.0 ** 2 / max(mse, 1e-30))


def _em_sheet(shape, seed, dtype=np.uint16):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:shape[-2], 0:shape[-1]]
    base = 90 + 60 * np.sin(yy / 37.0) * np.cos(xx / 53.0)
    return rng.poisson(np.broadcast_to(base, shape)).clip(0, 255).astype(dtype)


def test_config3_rdresunet_sliding_4096_sheet():
    """RDResUNet scale 4 on a 4096^2 uint16 sheet, SlidingDataset(hr_res=512, overlap=128) -> 100 tiles -> 3968^2 sheet."""
    from oracle.models import rdresunet_forward
    from pssr2_b200.data import SlidingDataset
    from pssr2_b200.models import RDResUNet
    from pssr2_b200.predict import predict_images
    from pssr2_b200.util import reassemble_sheets
    torch.manual_seed(0)
    model = RDResUNet().eval()
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    sheet = _em_sheet((1, 4096, 4096), 3)
    ds = SlidingDataset({"sheet": sheet}, hr_res=512, lr_scale=4, overlap=128, val_split=1, crappifier=None)
    assert len(ds) == 100                                           # (4096-512)//384+1 = 10 per axis, remainder dropped
    preds = predict_images(model, ds, device="cuda", batch_size=50, out_dir=None)
    assert len(preds) == 100
    # two tiles (a corner and an interior one) against the fp32 oracle chained like the reference
    for i in (0, 57):
        hr = OP.sliding_window(sheet, 512, 384, None, 1, i)
        _, lr = OP.gen_pair(hr, 512, 4, None)
        want = OP.pred_array(rdresunet_forward(sd, torch.as_tensor(lr)[None]).numpy())[0]
        got = preds[f"sheet_{i}_0"]
        d = np.abs(got.astype(int) - want.astype(int))
        assert got.shape == (1, 512, 512) and d.max() <= 1 and (d != 0).mean() < 0.02 and _psnr(got, want) >= 50.0, (d.max(), (d != 0).mean())
    # the stitched sheet is bit-exact the reference's overlap-average of the SAME tiles
    sheets = reassemble_sheets(preds, ds, lr_scale=1, overlap=128, margin=32, out_dir=None)
    tiles = np.stack([preds[f"sheet_{i}_0"][0] for i in range(100)])
    assert sheets[0].shape == (1, 3968, 3968)
    assert np.array_equal(sheets[0], OP.stitch_sheets(tiles, 10, 10, 128, 32))
    # partition invariance: another batch size gives the same tiles bit for bit
    again = predict_images(model, ds, device="cuda", batch_size=7, out_dir=None)
    assert all(np.array_equal(again[k], preds[k]) for k in preds)


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16])
def test_config4_crappifier_only_2048_tiles(dtype):
    """Poisson + AdditiveGaussian downscale-crappify of 2048^2 tiles -> 512^2: bit-exact with injected draws; the Philox stream
    does not depend on how the tiles are partitioned into launches."""
    from pssr2_b200 import ops
    rng = np.random.default_rng(11)
    n = 3
    tiles = rng.integers(0, 256 if dtype == np.uint8 else 600, (n, 2048, 2048)).astype(dtype)
    dev = torch.as_tensor(tiles.view(np.int16) if dtype == np.uint16 else tiles).cuda()
    table = ops.TileTable([dev], [0] * n, list(range(n)), [0] * n, [0] * n, [2048] * n, [2048] * n)
    y = rng.poisson(60, (n, 1, 512, 512)).astype(np.int64)
    g = rng.normal(0, 13, (n, 1, 512, 512))
    specs = [ops.NoiseSpec(1, 1, 0, True, torch.as_tensor(y).cuda()), ops.NoiseSpec(2, 13, 0, True, torch.as_tensor(g).cuda())]
    lr, _, _ = ops.crappify(table, 2048, 4, specs, clip_between=True)
    for i in range(n):
        _, want = OP.gen_pair(tiles[i:i + 1], 2048, 4, [("poisson", y[i], 1, 0), ("gaussian", g[i])])
        assert np.array_equal(lr[i].cpu().numpy(), want), f"tile {i} differs from the oracle"
    free = [ops.NoiseSpec(1, 1, 0), ops.NoiseSpec(2, 13, 0)]
    whole, _, _ = ops.crappify(table, 2048, 4, free, clip_between=True, seed=5, tile_index0=100)
    t1 = ops.TileTable([dev], [0], [2], [0], [0], [2048], [2048])
    part, _, _ = ops.crappify(t1, 2048, 4, free, clip_between=True, seed=5, tile_index0=102)
    assert torch.equal(whole[2], part[0])
    assert float(whole.min()) >= 0 and float(whole.max()) <= 255 and bool((whole == whole.round()).all())


def test_config5_scale8_multiframe_metrics():
    """ResUNet(channels=[5,1], scale=8): 5 x 2048^2 uint16 -> LR 5 x 256^2 -> 2048^2, test_metrics PSNR/SSIM."""
    from oracle.models import resunet_forward
    from pssr2_b200.data import ImageDataset
    from pssr2_b200.models import ResUNet
    from pssr2_b200.predict import predict_images, test_metrics as run_metrics
    torch.manual_seed(0)
    model = ResUNet(channels=[5, 1], scale=8).eval()
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    stacks = [_em_sheet((5, 2048, 2048), 20 + k) for k in range(2)]
    ds = ImageDataset(stacks, hr_res=2048, lr_scale=8, n_frames=[5, 1], val_split=1, crappifier=None)
    assert len(ds) == 2
    preds = predict_images(model, ds, device="cuda", batch_size=2, out_dir=None)
    names = sorted(preds)
    hr, lr = OP.gen_pair(stacks[0], 2048, 8, None, n_frames=[5, 1])
    assert lr.shape == (5, 256, 256) and hr.shape == (1, 2048, 2048)
    want = OP.pred_array(resunet_forward(sd, torch.as_tensor(lr)[None]).numpy())[0]
    got = preds[names[0]]
    d = np.abs(got.astype(int) - want.astype(int))
    assert got.shape == (1, 2048, 2048) and d.max() <= 1 and (d != 0).mean() < 0.02 and _psnr(got, want) >= 50.0, (d.max(), (d != 0).mean())
    # metrics of the SAME prediction: device kernels vs the oracle's formulas at full size
    m = run_metrics(model, ds, device="cuda", norm=False, avg=False, item0_quirk=False, batch_size=2)
    a = OP.pred_array(hr[None])
    mse, pixel, psnr, ssim = OP.image_metrics(a[0], got)
    assert abs(m["mse"][0] - mse) <= 1e-3 * mse + 1e-9 and abs(m["psnr"][0] - psnr) <= 1e-3 and abs(m["ssim"][0] - ssim) <= 1e-3
    assert abs(m["pixel"][0] - pixel) <= 1e-3 * pixel + 1e-6
