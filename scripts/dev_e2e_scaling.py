"""Developer probe: fixed vs per-step cost of the e2e predict_images call (not a bench)."""
import sys, time
import torch
sys.path.insert(0, ".")
from bench import _synthetic_tiles, BATCH, TILE, SCALE
from pssr2_b200.crappifiers import AdditiveGaussian, MultiCrappifier, Poisson
from pssr2_b200.data import ImageDataset
from pssr2_b200.models import ResUNet
from pssr2_b200.predict import predict_images
import contextlib, io
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = ResUNet().eval(); model.precision = "fp16"; model = model.to(dev)
crap = MultiCrappifier(Poisson(), AdditiveGaussian())
host = [_synthetic_tiles(BATCH, i, dev).cpu().pin_memory() for i in range(2)]
def run(n):
    stacks = [host[i % 2] for i in range(n)]
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ds = ImageDataset(stacks, hr_res=TILE, lr_scale=SCALE, crappifier=crap, n_frames=1, val_split=1, device=dev)
    t1 = time.perf_counter()
    with contextlib.redirect_stderr(io.StringIO()):
        p = predict_images(model, ds, device=str(dev), batch_size=BATCH, out_dir=None)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    return (t1 - t0) * 1e3, (t2 - t1) * 1e3, p
for n in (4, 10, 20, 40, 10, 20, 40):
    for rep in range(2):
        a, b, p = run(n)
        del p
    print(f"n={n:3d}: dataset ctor {a:7.2f} ms, predict {b:7.2f} ms -> {(a+b)/n:.3f} ms/step, {n*BATCH*TILE*TILE/(a+b)/1e3:.1f} HR MP/s")
