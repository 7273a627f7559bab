"""Developer timing of the HBM-bound ops (crappify / metrics / normalize / stitch) with CUDA events."""
import sys
sys.path.insert(0, ".")
import torch
from pssr2_b200 import ops
from bench import _synthetic_tiles

def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

B, T = 64, 512
dev = torch.device("cuda")
b16 = _synthetic_tiles(B, 1, dev)
b8 = b16.to(torch.uint8)
for name, sheet in (("u16", b16), ("u8", b8)):
    table = ops.TileTable([sheet], [0] * B, list(range(B)), [0] * B, [0] * B, [T] * B, [T] * B)
    eb = sheet.element_size()
    for label, specs in (("no noise", None), ("poisson+gauss philox", [ops.NoiseSpec(1, 1, 0), ops.NoiseSpec(2, 13, 0)])):
        t = timeit(lambda: ops.crappify(table, T, 4, specs, clip_between=True))
        byts = B * T * T * eb + B * 128 * 128 * 4
        print(f"crappify {name} {label:22s}: {t*1000:8.1f} us  {byts/t/1e6:8.1f} GB/s")
    t = timeit(lambda: ops.crappify(table, T, 4, None, want_lr=False, want_hr_u8=True))
    print(f"hr_gather u8 out {name}: {t*1000:8.1f} us  {(B*T*T*(eb+1))/t/1e6:8.1f} GB/s")
a = b8
b = (b8.float() * 0.9 + 5).to(torch.uint8)
t = timeit(lambda: ops.metric_sums(a, b))
print(f"metric_sums 64x512^2: {t*1000:8.1f} us  {2*B*T*T/t/1e6:8.1f} GB/s")
t = timeit(lambda: ops.metric_sums(a, b, want_ssim=False))
print(f"metric_sums (no ssim buf): {t*1000:8.1f} us")
t = timeit(lambda: ops.normalize_preds_u8(a, b))
print(f"normalize_preds 64x512^2: {t*1000:8.1f} us  {4*B*T*T/t/1e6:8.1f} GB/s (2 reads + 2 writes)")
tiles = torch.randint(0, 256, (100, 512, 512), dtype=torch.uint8, device=dev)
t = timeit(lambda: ops.stitch(tiles, 10, 10, 128, 32))
print(f"stitch 100x512^2 -> 3968^2: {t*1000:8.1f} us  {(100*512*512 + 3968*3968)/t/1e6:8.1f} GB/s")
big = _synthetic_tiles(16, 2, dev)  # stand-in
