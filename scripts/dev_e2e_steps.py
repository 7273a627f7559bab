"""Developer probe: e2e ms/step as a function of the number of steps, with and without the dataset upload in the timed region."""
import sys, time, contextlib, io
import torch
sys.path.insert(0, ".")
from bench import _synthetic_tiles, BATCH, TILE, SCALE
from pssr2_b200.crappifiers import AdditiveGaussian, MultiCrappifier, Poisson
from pssr2_b200.data import ImageDataset
from pssr2_b200.models import ResUNet
from pssr2_b200.predict import predict_images

dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = ResUNet().eval().to(dev)
crap = MultiCrappifier(Poisson(), AdditiveGaussian())
host = [_synthetic_tiles(BATCH, s, dev).cpu().pin_memory() for s in (1, 2)]

def mk(steps):
    ds = ImageDataset([host[i % 2] for i in range(steps)], hr_res=TILE, lr_scale=SCALE, crappifier=crap, n_frames=1, val_split=1, device=dev)
    ds.rank_local = True
    return ds

def T():
    torch.cuda.synchronize(); return time.perf_counter()

with contextlib.redirect_stderr(io.StringIO()):
    for steps in (12, 12, 25, 50, 50, 100):
        t0 = T(); ds = mk(steps); p = predict_images(model, ds, device=str(dev), batch_size=BATCH, out_dir=None); t1 = T()
        del p
        ds2 = mk(steps); t2 = T(); p = predict_images(model, ds2, device=str(dev), batch_size=BATCH, out_dir=None); t3 = T()
        del p
        t4 = T(); p = predict_images(model, ds2, device=str(dev), batch_size=BATCH, out_dir=None, keep_on_device=True); t5 = T()
        del p
        print(f"steps {steps:3d}: full e2e {1e3*(t1-t0)/steps:.3f} ms/step | ctor outside {1e3*(t3-t2)/steps:.3f} | resident sheets + keep_on_device {1e3*(t5-t4)/steps:.3f}", flush=True)
