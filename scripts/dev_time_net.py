"""Developer timing of one ResUNet plan: whole forward and per-op CUDA-event times (not a bench)."""
import sys, time
import torch
sys.path.insert(0, ".")
from pssr2_b200.models import ResUNet

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
prec = sys.argv[2] if len(sys.argv) > 2 else "fp16"
torch.manual_seed(0)
m = ResUNet().eval()
m.precision = prec
m = m.cuda()
x = torch.randint(0, 256, (B, 1, 128, 128), device="cuda").float()
st, _ = m._state(x)
plan = st["plan"]
for _ in range(3):
    plan.run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 10
e0.record()
for _ in range(reps):
    plan.run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
alg = 63.305e9 * B
print(f"B={B} {prec}: forward {ms:.3f} ms  -> {B*512*512/ms/1e3:.1f} HR MP/s, {alg/ms/1e9:.1f} TFLOP/s algorithmic, ops={len(plan)}")
tot = 0
for i, (kind, r) in enumerate(plan.records):
    for _ in range(2):
        plan.run(i, 1)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        plan.run(i, 1)
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 5
    tot += t
    extra = ""
    if kind == "conv":
        segs = r["segs"]
        extra = f"n={r['n']} {r['Ho']}x{r['Wo']} kb={sum(sg[1]*sg[2] for sg in segs)} shuffle={r['shuffle']} issued {r['issued_flops']/t/1e9:.0f} TF/s"
    print(f"  op{i:02d} {kind:8s} {t*1000:8.1f} us  {extra}")
print(f"sum of ops {tot:.3f} ms")
