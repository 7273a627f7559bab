"""Developer timeline of one conv layer (PSSR_DBG=16): per-CTA clock64 stamps of the MMA / epilogue warps (not a bench)."""
import os, sys
os.environ["PSSR_DBG"] = str(int(os.environ.get("PSSR_DBG", "0")) | 16)
import numpy as np, torch
sys.path.insert(0, ".")
from pssr2_b200 import _lib
from scripts.dev_power_probe_lib import make

cfgs = [(64, 128, 128, 64, 64), (64, 64, 64, 128, 128), (64, 32, 32, 256, 256)]
for ci in [int(a) for a in sys.argv[1:]] or [2]:
    cfg = cfgs[ci]
    plan, keep = make(*cfg)
    for _ in range(3):
        plan.run()
    torch.cuda.synchronize()
    buf = np.zeros(148 * 256, dtype=np.int64)
    rc = _lib.lib().pssr_debug_trace(buf.ctypes.data, buf.size)
    assert rc == 0
    tr = buf.reshape(148, 256)
    print("DBG", os.environ["PSSR_DBG"], "cfg", cfg)
    st = tr[0][128:256]; st = st[st > 0]
    print("  cta 0 unit 1 per-stage b_full-wait-done deltas:", np.diff(st)[:60].tolist())
    for cta in (0, 73):
        t = tr[cta]; t0 = t[0]
        units = [u for u in range(30) if t[2 + 2 * u] > 0 and t[2 + 2 * u] >= t0]
        print(f" cta {cta}: setup {t[1]-t0} cyc; exit at {t[127]-t0}; units {len(units)}")
        for u in units:
            print(f"   unit {u}: mma first-issue {t[2+2*u]-t0:7d} commit-issued {t[3+2*u]-t0:7d} | acc ready {t[64+2*u]-t0:7d} epi done {t[65+2*u]-t0:7d}  (epi {t[65+2*u]-t[64+2*u]})")
