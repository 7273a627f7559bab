"""Developer probe: device timeline of the bench's e2e leg (torch.profiler / CUPTI): where the compute stream idles."""
import sys
import time
import torch
from torch.profiler import ProfilerActivity, profile
sys.path.insert(0, ".")
from bench import _synthetic_tiles, BATCH, TILE, SCALE
from pssr2_b200.crappifiers import AdditiveGaussian, MultiCrappifier, Poisson
from pssr2_b200.data import ImageDataset
from pssr2_b200.models import ResUNet
from pssr2_b200.predict import predict_images

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 12
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


for rep in range(2):
    p = run(); del p
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    t0 = time.perf_counter(); p = run(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"profiled run: {1e3*dt/steps:.3f} ms/step")
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
print("device events:", len(evs))
t_first = evs[0].time_range.start
kern = [e for e in evs if "Memcpy" not in e.name and "Memset" not in e.name]
cpy = [e for e in evs if "Memcpy" in e.name]
print(f"span {(evs[-1].time_range.end - t_first)/1e3:.2f} ms, kernels busy {sum(e.time_range.end - e.time_range.start for e in kern)/1e3:.2f} ms, "
      f"copies {sum(e.time_range.end - e.time_range.start for e in cpy)/1e3:.2f} ms over {len(cpy)} copies")
gaps = []
for a, b in zip(kern[:-1], kern[1:]):
    g = b.time_range.start - a.time_range.end
    if g > 15:
        gaps.append((g, a.name[:50], b.name[:50], (a.time_range.end - t_first) / 1e3))
print(f"gaps > 15 us between consecutive kernels: {len(gaps)}, total {sum(g[0] for g in gaps)/1e3:.2f} ms")
for g in sorted(gaps, reverse=True)[:25]:
    print(f"  {g[0]:8.1f} us at {g[3]:8.2f} ms   {g[1]}  ->  {g[2]}")
# per-kernel-name totals
tot = {}
for e in evs:
    k = e.name[:60]
    tot.setdefault(k, [0, 0.0])
    tot[k][0] += 1
    tot[k][1] += e.time_range.end - e.time_range.start
for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"  {t/1e3/steps:8.3f} ms/step  x{n/steps:5.1f}  {k}")
for e in cpy[:2 * 6]:
    print(f"  copy {e.name[:30]} {(e.time_range.start - t_first)/1e3:8.2f} ms  dur {(e.time_range.end - e.time_range.start):8.1f} us")
cr = [e for e in kern if "crappify" in e.name]
print("step periods (ms) from crappify start to crappify start:")
print("  " + " ".join(f"{(b.time_range.start - a.time_range.start)/1e3:.2f}" for a, b in zip(cr[:-1], cr[1:])))
print("  first crappify at %.2f ms, last kernel end %.2f ms, last copy end %.2f ms" % ((cr[0].time_range.start - t_first) / 1e3,
      (kern[-1].time_range.end - t_first) / 1e3, (cpy[-1].time_range.end - t_first) / 1e3))
h2d = [e for e in cpy if "HtoD" in e.name and e.time_range.end - e.time_range.start > 100]
print("  bulk H2D: first start %.2f ms, last end %.2f ms" % ((h2d[0].time_range.start - t_first) / 1e3, (h2d[-1].time_range.end - t_first) / 1e3))
