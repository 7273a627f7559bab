"""Developer probe: host-side profile of the bench's e2e leg (ImageDataset of pinned stacks + ONE predict_images call)."""
import cProfile
import pstats
import sys
import time
import torch
sys.path.insert(0, ".")
from bench import _synthetic_tiles, BATCH, TILE, SCALE
from pssr2_b200.crappifiers import AdditiveGaussian, MultiCrappifier, Poisson
from pssr2_b200.data import ImageDataset
from pssr2_b200.models import ResUNet
from pssr2_b200.predict import predict_images

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = ResUNet().eval().to(dev)
crap = MultiCrappifier(Poisson(), AdditiveGaussian())
host = [_synthetic_tiles(BATCH, s, dev).cpu().pin_memory() for s in (1, 2)]
stacks = [host[i % 2] for i in range(steps)]


def run():
    ds = ImageDataset(list(stacks), hr_res=TILE, lr_scale=SCALE, crappifier=crap, n_frames=1, val_split=1, device=dev)
    ds.rank_local = True
    return predict_images(model, ds, device=str(dev), batch_size=BATCH, out_dir=None)


for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    p = run()
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"rep {rep}: {1e3*dt/steps:.3f} ms/step  {steps*BATCH*TILE*TILE/dt/1e6:.0f} HR MP/s", flush=True)
    del p
pr = cProfile.Profile()
torch.cuda.synchronize(); t0 = time.perf_counter()
pr.enable(); p = run(); torch.cuda.synchronize(); pr.disable()
print(f"profiled: {1e3*(time.perf_counter()-t0)/steps:.3f} ms/step")
pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
