"""Developer timeline (PSSR_DBG=16) of one op of the ResUNet plan (not a bench)."""
import os, sys
os.environ["PSSR_DBG"] = str(int(os.environ.get("PSSR_DBG", "0")) | 16)
import numpy as np, torch
sys.path.insert(0, ".")
from pssr2_b200 import _lib
from pssr2_b200.models import ResUNet

op = int(sys.argv[1]) if len(sys.argv) > 1 else 41
B = 64
torch.manual_seed(0)
m = ResUNet().eval()
m.precision = os.environ.get("PREC", "fp16")
m = m.cuda()
x = torch.randint(0, 256, (B, 1, 128, 128), device="cuda").float()
st, _ = m._state(x)
plan = st["plan"]
for _ in range(2):
    plan.run()
torch.cuda.synchronize()
plan.run(op, 1)
torch.cuda.synchronize()
buf = np.zeros(148 * 256, dtype=np.int64)
assert _lib.lib().pssr_debug_trace(buf.ctypes.data, buf.size) == 0
tr = buf.reshape(148, 256)
print("op", op, plan.records[op][0], {k: v for k, v in plan.records[op][1].items() if k in ("n", "Ho", "Wo", "segs", "shuffle")})
for cta in (0, 1):
    t = tr[cta]; t0 = t[0]
    print(f" cta {cta}: setup {t[1]-t0} cyc; exit at {t[127]-t0}")
    for u in range(12):
        if t[64 + 2 * u] <= 0:
            continue
        m_ = f"buffer free {t[2+2*u]-t0:7d} A landed {t[128+2*u]-t0:7d} committed {t[3+2*u]-t0:7d} tail issued {t[192+u]-t0:7d}" if t[2 + 2 * u] > 0 else " " * 80
        extra = f" | tail wait from {t[129+2*u]-t0:7d}" if t[129 + 2 * u] > 0 else ""
        if u < 8 and t[232 + 2 * u] > 0:
            extra += f" handed over {t[232+2*u]-t0:7d} z seen {t[233+2*u]-t0:7d}"
        print(f"   unit {u}: {m_} | acc ready {t[64+2*u]-t0:7d} epi done {t[65+2*u]-t0:7d} (epi {t[65+2*u]-t[64+2*u]}){extra}")
    if t[244] > 0:
        names = ["accumulator buffer", "input rows", "weight stages", "tail activations"]
        print("   MMA warp waits (cycles, whole launch): " + ", ".join(f"{n} {t[240+k]} ({100*t[240+k]/t[244]:.1f} %)" for k, n in enumerate(names)) + f"; issue loop {t[244]}")
