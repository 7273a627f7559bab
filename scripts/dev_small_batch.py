"""Developer probe: predict_images throughput at small batch sizes (host overhead per batch), not a bench."""
import sys, time, contextlib, io, cProfile, pstats
import torch
sys.path.insert(0, ".")
from bench import _synthetic_tiles, TILE, SCALE
from pssr2_b200.crappifiers import AdditiveGaussian, MultiCrappifier, Poisson
from pssr2_b200.data import ImageDataset
from pssr2_b200.models import ResUNet
from pssr2_b200.predict import predict_images

dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = ResUNet().eval().to(dev)
crap = MultiCrappifier(Poisson(), AdditiveGaussian())
host = _synthetic_tiles(256, 1, dev).cpu().pin_memory()
for bs in (1, 4, 16, 64):
    with contextlib.redirect_stderr(io.StringIO()):
        for rep in range(2):
            ds = ImageDataset([host], hr_res=TILE, lr_scale=SCALE, crappifier=crap, n_frames=1, val_split=1, device=dev)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            p = predict_images(model, ds, device=str(dev), batch_size=bs, out_dir=None)
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"batch_size {bs:3d}: {1e3*dt/(256/bs):7.3f} ms per batch, {256*TILE*TILE/dt/1e6:7.0f} HR MP/s", flush=True)
if len(sys.argv) > 1:
    ds = ImageDataset([host], hr_res=TILE, lr_scale=SCALE, crappifier=crap, n_frames=1, val_split=1, device=dev)
    pr = cProfile.Profile(); pr.enable()
    with contextlib.redirect_stderr(io.StringIO()):
        p = predict_images(model, ds, device=str(dev), batch_size=1, out_dir=None)
    torch.cuda.synchronize(); pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(22)
