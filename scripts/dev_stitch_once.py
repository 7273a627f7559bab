"""Developer run for ncu: one stitch launch on 8 sheets of 64 tiles (not a bench)."""
import sys
import torch
sys.path.insert(0, ".")
from pssr2_b200 import ops
t = torch.randint(0, 256, (512, 512, 512), dtype=torch.uint8, device="cuda")
for _ in range(2):
    ops.stitch(t, 8, 8, 128, 32)
torch.cuda.synchronize()
torch.cuda.profiler.start()
ops.stitch(t, 8, 8, 128, 32)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
