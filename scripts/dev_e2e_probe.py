"""Developer probe: where the host-side time of the e2e predict_images call goes (not a bench)."""
import sys, time
import torch
sys.path.insert(0, ".")
from bench import _synthetic_tiles, BATCH, TILE, SCALE
from pssr2_b200.crappifiers import AdditiveGaussian, MultiCrappifier, Poisson
from pssr2_b200.data import ImageDataset
from pssr2_b200.models import ResUNet
from pssr2_b200.predict import predict_images, _batch, _pred_u8

dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = ResUNet().eval(); model.precision = "fp16"; model = model.to(dev)
crap = MultiCrappifier(Poisson(), AdditiveGaussian())
host = _synthetic_tiles(BATCH, 1, dev).cpu().pin_memory()

def T():
    torch.cuda.synchronize(); return time.perf_counter()

for rep in range(4):
    t0 = T()
    ds = ImageDataset([host], hr_res=TILE, lr_scale=SCALE, crappifier=crap, n_frames=1, val_split=1, device=dev)
    t1 = T()
    idxs = list(ds.val_idx)[:BATCH]
    lr, hr8 = _batch(ds, idxs, str(dev), want_hr_u8=False)
    t2 = T()
    out8 = _pred_u8(model, lr)
    t3 = T()
    arr = out8[:, :, :TILE, :TILE].contiguous().cpu().numpy()
    t4 = T()
    outs = {ds._get_name(i): arr[i] for i in range(BATCH)}
    t5 = T()
    print(f"rep {rep}: dataset ctor+H2D {1e3*(t1-t0):.2f} ms | batch (table + crappify) {1e3*(t2-t1):.2f} | forward {1e3*(t3-t2):.2f} | D2H pageable {1e3*(t4-t3):.2f} | dict {1e3*(t5-t4):.2f}")
for rep in range(3):
    t0 = T()
    ds = ImageDataset([host], hr_res=TILE, lr_scale=SCALE, crappifier=crap, n_frames=1, val_split=1, device=dev)
    p = predict_images(model, ds, device=str(dev), batch_size=BATCH, out_dir=None)
    t1 = T()
    print(f"predict_images call {1e3*(t1-t0):.2f} ms")
pin = torch.empty(BATCH, 1, TILE, TILE, dtype=torch.uint8).pin_memory()
t0 = T(); pin.copy_(out8, non_blocking=True); t1 = T()
print(f"D2H pinned {1e3*(t1-t0):.2f} ms")
t0 = T(); x = torch.empty(BATCH, 1, TILE, TILE, dtype=torch.uint8, pin_memory=True); t1 = T()
print(f"pinned alloc {1e3*(t1-t0):.2f} ms")
del x
t0 = T(); x = torch.empty(BATCH, 1, TILE, TILE, dtype=torch.uint8, pin_memory=True); t1 = T()
print(f"pinned alloc (cached) {1e3*(t1-t0):.2f} ms")
