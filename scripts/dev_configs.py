"""Developer run of BASELINE.json configs 3, 4, 5 (timings with CUDA events; parity spot checks)."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import torch
from pssr2_b200 import ops
from pssr2_b200.crappifiers import AdditiveGaussian, MultiCrappifier, Poisson
from pssr2_b200.data import SlidingDataset, ImageDataset
from pssr2_b200.models import ResUNet, RDResUNet
from pssr2_b200.predict import predict_images, test_metrics
from pssr2_b200.util import reassemble_sheets

def ev_time(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

which = sys.argv[1:] or ["4", "5", "3"]
dev = torch.device("cuda")
rng = np.random.default_rng(0)

if "4" in which:
    print("== config 4: crappifier-only, Poisson+AdditiveGaussian, 2048^2 -> 512^2 tiles")
    for dtype, nt in ((torch.uint8, 512), (torch.int16, 512)):
        sheet = torch.randint(0, 256, (nt, 2048, 2048), device=dev, dtype=torch.int16 if dtype == torch.int16 else torch.uint8)
        table = ops.TileTable([sheet], [0] * nt, list(range(nt)), [0] * nt, [0] * nt, [2048] * nt, [2048] * nt)
        specs = [ops.NoiseSpec(1, 1, 0), ops.NoiseSpec(2, 13, 0)]
        eb = sheet.element_size()
        for label, sp in (("philox noise", specs), ("no noise", None)):
            t = ev_time(lambda: ops.crappify(table, 2048, 4, sp, clip_between=True, seed=3))
            byts = nt * (2048 * 2048 * eb + 512 * 512 * 4)
            print(f"  {'u16' if eb == 2 else 'u8 '} {nt} tiles {label:13s}: {t:8.3f} ms  {byts / t / 1e6:8.1f} GB/s  ({nt * 2048 * 2048 / t / 1e9:.2f} tera-HR-px/s)")
        del sheet, table
        torch.cuda.empty_cache()

if "5" in which:
    print("== config 5: ResUNet(channels=[5,1], scale=8), 5x2048^2 uint16 -> LR 5x256^2 -> 2048^2, test_metrics")
    torch.manual_seed(0)
    model = ResUNet(channels=[5, 1], scale=8).eval().to(dev)
    B = 8
    x = torch.randint(0, 256, (B, 5, 256, 256), device=dev).float()
    st, _ = model._state(x)
    t = ev_time(lambda: st["plan"].run(), reps=3)
    print(f"  forward B={B}: {t:.3f} ms -> {B * 2048 * 2048 / t / 1e3:.1f} HR MP/s, {512.06e9 * B / t / 1e9:.1f} TFLOP/s algorithmic")
    for i, (kind, r) in enumerate(st["plan"].records):
        tt = ev_time(lambda: st["plan"].run(i, 1), reps=2, warm=1)
        if tt > 0.3: print(f"     op{i:02d} {kind:8s} {tt:8.3f} ms" + (f"  n={r['n']} {r['Ho']}x{r['Wo']}" if kind == "conv" else ""))
    # parity spot check vs oracle on one tile
    from oracle.models import resunet_forward
    xs = x[:1].cpu()
    want = resunet_forward({k: v.cpu() for k, v in model.state_dict().items()}, xs)
    got = model(x[:1]).cpu()
    print(f"  parity vs fp32 oracle: max-abs {float((got - want).abs().max()):.4f}")
    imgs = rng.poisson(90, (5, 2048, 2048)).clip(0, 255).astype(np.uint16)
    ds = ImageDataset([imgs], hr_res=2048, lr_scale=8, n_frames=[5, 1], val_split=1, crappifier=MultiCrappifier(Poisson(), AdditiveGaussian()))
    m = test_metrics(model, ds, device="cuda")
    print("  test_metrics:", m)
    del model, st
    torch.cuda.empty_cache()

if "3" in which:
    print("== config 3: RDResUNet scale 4 on a 4096^2 uint16 sheet, SlidingDataset(overlap=128) -> 100 tiles -> stitch 3968^2")
    torch.manual_seed(0)
    model = RDResUNet().eval().to(dev)
    sheet = rng.poisson(90, (1, 4096, 4096)).clip(0, 255).astype(np.uint16)
    ds = SlidingDataset({"sheet": sheet}, hr_res=512, lr_scale=4, overlap=128, val_split=1, crappifier=MultiCrappifier(Poisson(), AdditiveGaussian()))
    print("  tiles:", len(ds))
    x = torch.randint(0, 256, (50, 1, 128, 128), device=dev).float()
    st, _ = model._state(x)
    t = ev_time(lambda: st["plan"].run(), reps=3)
    print(f"  forward B=50: {t:.3f} ms -> {50 * 512 * 512 / t / 1e3:.1f} HR MP/s, {107.15e9 * 50 / t / 1e9:.1f} TFLOP/s algorithmic, ops={len(st['plan'])}")
    agg = {}
    for i, (kind, r) in enumerate(st["plan"].records):
        tt = ev_time(lambda: st["plan"].run(i, 1), reps=2, warm=1)
        agg[kind] = agg.get(kind, 0) + tt
    print("  per-kind ms:", {k: round(v, 3) for k, v in agg.items()})
    t0 = time.perf_counter()
    preds = predict_images(model, ds, device="cuda", batch_size=50, out_dir=None)
    sheets = reassemble_sheets(preds, ds, lr_scale=1, overlap=128, margin=32, out_dir=None)
    torch.cuda.synchronize()
    print(f"  predict_images+reassemble (host dict path) {time.perf_counter() - t0:.3f} s, sheet {sheets[0].shape}")
