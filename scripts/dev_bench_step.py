import sys, time
sys.path.insert(0, ".")
import torch
from pssr2_b200 import ops
from pssr2_b200.crappifiers import AdditiveGaussian, MultiCrappifier, Poisson
from pssr2_b200.models import ResUNet
from bench import _synthetic_tiles, TILE, SCALE, BATCH
dev = torch.device("cuda")
torch.manual_seed(0)
model = ResUNet().eval().to(dev)
crap = MultiCrappifier(Poisson(), AdditiveGaussian())
b = _synthetic_tiles(BATCH, 1, dev)
table = ops.TileTable([b], [0] * BATCH, list(range(BATCH)), [0] * BATCH, [0] * BATCH, [TILE] * BATCH, [TILE] * BATCH)
specs = crap.noise_specs()
part = torch.zeros(3, dtype=torch.float64, device=dev)
def T(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
lr, _, hr8 = ops.crappify(table, TILE, SCALE, specs, clip_between=True, seed=1, want_hr_u8=True)
_, out8 = model.forward_u8(lr)
print("crappify      ms", T(lambda: ops.crappify(table, TILE, SCALE, specs, clip_between=True, seed=1, want_hr_u8=True)))
print("forward_u8    ms", T(lambda: model.forward_u8(lr)))
print("metric_sums   ms", T(lambda: ops.metric_sums(hr8[:, 0], out8[:, 0])))
sq, ss = ops.metric_sums(hr8[:, 0], out8[:, 0])
def red():
    part[0] = sq.sum(); part[1] = ss.sum(); part[2] = BATCH
print("reductions    ms", T(red))
def full(i=[0]):
    lr, _, hr8 = ops.crappify(table, TILE, SCALE, specs, clip_between=True, seed=i[0], want_hr_u8=True)
    _, out8 = model.forward_u8(lr)
    sq, ss = ops.metric_sums(hr8[:, 0], out8[:, 0])
    part[0] = sq.sum(); part[1] = ss.sum(); part[2] = BATCH
    i[0] += 1
print("full step     ms", T(full, 20))
