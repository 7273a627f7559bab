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
    print("DBG", os.environ["PSSR_DBG"], "NOPAIR", os.environ.get("PSSR_V3_NOPAIR"), "cfg", cfg)
    t = tr[0]
    print("  cta0 warp4 unit2 epilogue: ready->idx %d, ->ld0 %d, ->st0 %d, ->ld1 %d, ->st1 %d, ->loopend %d, ->fence %d, ->arrived %d" % (
        t[224]-t[64+4], t[225]-t[224], t[226]-t[225], t[227]-t[226], t[228]-t[227], t[230]-t[228], t[231]-t[230], t[65+4]-t[231]))
    for cta in (0, 1):
        t = tr[cta]; t0 = t[0]
        print(f" cta {cta}: setup {t[1]-t0} cyc; exit at {t[127]-t0}")
        for u in range(8):
            if t[64 + 2 * u] <= 0:
                continue
            m = f"buffer free {t[2+2*u]-t0:7d} A landed {t[128+2*u]-t0:7d} committed {t[3+2*u]-t0:7d}" if t[2 + 2 * u] > 0 else " " * 60
            print(f"   unit {u}: {m} | acc ready {t[64+2*u]-t0:7d} epi done {t[65+2*u]-t0:7d} (epi {t[65+2*u]-t[64+2*u]})")
