"""Developer probe: does capturing the plan's launches in a CUDA graph shorten the forward? (not a bench)"""
import sys
import torch
sys.path.insert(0, ".")
from pssr2_b200.models import ResUNet
torch.manual_seed(0)
m = ResUNet().eval().cuda()
x = torch.randint(0, 256, (64, 1, 128, 128), device="cuda").float()
st, _ = m._state(x)
plan = st["plan"]
def timeit(fn, reps=20):
    for _ in range(3): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
print("direct launches: %.3f ms" % timeit(plan.run))
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    plan.run(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        plan.run()
    torch.cuda.synchronize()
    print("graph replay   : %.3f ms" % timeit(g.replay))
