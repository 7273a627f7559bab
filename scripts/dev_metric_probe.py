import sys
sys.path.insert(0, ".")
import torch
from pssr2_b200 import ops
from bench import _synthetic_tiles
dev = torch.device("cuda")
def ev(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1000
a = _synthetic_tiles(64, 1, dev).to(torch.uint8)
cases = {"b=0.9a+5": (a.float() * 0.9 + 5).to(torch.uint8), "b=128 const": torch.full_like(a, 128), "b=a": a.clone(),
         "b=noise": torch.randint(0, 256, a.shape, device=dev, dtype=torch.uint8), "b=127/128 mix": (torch.rand(a.shape, device=dev) > 0.5).to(torch.uint8) + 127}
for k, b in cases.items():
    print(f"{k:16s} metric_sums {ev(lambda: ops.metric_sums(a, b)):8.1f} us")
