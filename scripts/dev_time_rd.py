import sys
sys.path.insert(0, ".")
import torch
from pssr2_b200.models import RDResUNet
B = int(sys.argv[1]) if len(sys.argv) > 1 else 50
torch.manual_seed(0)
m = RDResUNet().eval().cuda()
if len(sys.argv) > 2: m.precision = sys.argv[2]
x = torch.randint(0, 256, (B, 1, 128, 128), device="cuda").float()
st, _ = m._state(x)
plan = st["plan"]
for _ in range(3): plan.run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(3): plan.run()
torch.cuda.synchronize(); e0.record()
for _ in range(5): plan.run()
e1.record(); torch.cuda.synchronize()
print(f'forward {e0.elapsed_time(e1)/5:.3f} ms ({m.precision})')
rows = []
for i, (kind, r) in enumerate(plan.records):
    plan.run(i, 1); torch.cuda.synchronize()
    e0.record()
    for _ in range(3): plan.run(i, 1)
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 3
    extra = ""
    if kind == "conv":
        extra = f"n={r['n']:5d} nv={r['n_valid']:5d} {r['Ho']:3d}x{r['Wo']:3d} kb={sum(sg[1]*sg[2] for sg in r['segs']):4d} taps={[sg[1] for sg in r['segs']]} issued {r['issued_flops']/t/1e9:6.0f} TF/s"
    rows.append((t, i, kind, extra))
tot = sum(r[0] for r in rows)
print(f"B={B}: sum {tot:.3f} ms")
if len(sys.argv) > 3 and sys.argv[3] == "all":
    by = {}
    for t, i, kind, extra in rows: by[kind] = by.get(kind, 0) + t
    print("  per kind (ms):", {k: round(v, 3) for k, v in sorted(by.items(), key=lambda kv: -kv[1])})
    sel = sorted(rows, key=lambda r: r[1])
else:
    sel = sorted(rows, reverse=True)[:40]
for t, i, kind, extra in sel:
    print(f"  op{i:03d} {kind:8s} {t*1000:8.1f} us  {extra}")
