"""Developer probe (torchrun-aware): device timeline of config 3's predict_sheets call on every rank -- where the streams idle."""
import os, sys, time
import torch
from torch.profiler import ProfilerActivity, profile
sys.path.insert(0, ".")
import bench
from bench import _synthetic_tiles, _Placeholder, _quiet
from pssr2_b200.crappifiers import AdditiveGaussian, MultiCrappifier, Poisson
from pssr2_b200.data import SlidingDataset
from pssr2_b200.models import RDResUNet
from pssr2_b200.predict import predict_sheets

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl")
spr = int(sys.argv[1]) if len(sys.argv) > 1 else 4
torch.manual_seed(0)
model = RDResUNet().eval().to(dev)
srcs = {}
for s in range(spr * world):
    srcs[f"sheet{s}"] = (_synthetic_tiles(1, 4321 + s, dev, size=4096)[0:1].cpu().pin_memory() if s // spr == rank else _Placeholder((1, 4096, 4096)))
crap = MultiCrappifier(Poisson(), AdditiveGaussian())
mk = lambda: SlidingDataset(dict(srcs), hr_res=512, lr_scale=4, overlap=128, val_split=1, crappifier=crap, device=dev, preload=False)
with _quiet(), torch.no_grad():
    predict_sheets(model, mk(), device=str(dev), batch_size=50, margin=32)
    predict_sheets(model, mk(), device=str(dev), batch_size=50, margin=32)
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    stamps = []
    import pssr2_b200.predict as P
    t0 = time.perf_counter()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        out = predict_sheets(model, mk(), device=str(dev), batch_size=50, margin=32)
        torch.cuda.synchronize()
    dt = time.perf_counter() - t0
evs = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
t_first = evs[0].time_range.start
kern = [e for e in evs if "Memcpy" not in e.name and "Memset" not in e.name]
lines = [f"[rank {rank}] wall {dt*1e3:.1f} ms for {spr} sheets; device span {(evs[-1].time_range.end - t_first)/1e3:.1f} ms; first kernel at +? ; kernels {len(kern)}"]
gaps = []
for a, b in zip(kern[:-1], kern[1:]):
    g = b.time_range.start - a.time_range.end
    if g > 100:
        gaps.append((g, (a.time_range.end - t_first) / 1e3, a.name[:40], b.name[:40]))
lines.append(f"[rank {rank}] gaps > 100 us: {len(gaps)}, total {sum(g[0] for g in gaps)/1e3:.2f} ms")
for g in sorted(gaps, reverse=True)[:8]:
    lines.append(f"[rank {rank}]   {g[0]/1e3:7.2f} ms at {g[1]:8.2f} ms  {g[2]} -> {g[3]}")
for e in evs:
    if "Memcpy" in e.name and e.time_range.end - e.time_range.start > 200 or "nccl" in e.name.lower():
        lines.append(f"[rank {rank}]   {e.name[:44]:44s} at {(e.time_range.start - t_first)/1e3:8.2f} ms dur {(e.time_range.end - e.time_range.start)/1e3:7.2f} ms")
cpu = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CPU and e.time_range.end - e.time_range.start > 1500],
             key=lambda e: -(e.time_range.end - e.time_range.start))[:10]
for e in cpu:
    lines.append(f"[rank {rank}]   host op {e.name[:50]:50s} {(e.time_range.end - e.time_range.start)/1e3:7.2f} ms")
for r_ in range(world):
    if r_ == rank:
        print("\n".join(lines), flush=True)
    if world > 1: dist.barrier()
if world > 1: dist.destroy_process_group()
