"""Developer probe: run one conv layer back to back for a few seconds and sample SM clock / power (not a bench)."""
import sys, os, subprocess, time, threading
import torch
sys.path.insert(0, ".")
from pssr2_b200 import plan as P

def make(B, H, W, Cin, Cout, prec="fp16"):
    plan = P.Plan(prec)
    dt = plan.tdtype
    x = torch.randn(B, H, W, Cin, device="cuda").to(dt)
    w = torch.randn(Cout, Cin, 3, 3, device="cuda") / (3 * Cin ** 0.5)
    b = torch.zeros(Cout, device="cuda")
    wp = P.pack_weight([w], plan.dtype)
    out = torch.zeros(B, H, W, Cout, dtype=dt, device="cuda")
    plan.conv([P.View(x)], [(0, 9, P.ceil_div(Cin, 64))], wp, b, P.View(out), Ho=H, Wo=W, B=B, act=P.ACT_RELU)
    plan.finalize()
    return plan, (x, w, b, wp, out)

samples = []
stop = False
def sampler():
    while not stop:
        r = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-i", "0"], capture_output=True, text=True)
        try:
            c, pw = r.stdout.strip().split(",")
            samples.append((time.time(), float(c), float(pw)))
        except Exception:
            pass
        time.sleep(0.05)

cfg = [(64, 128, 128, 64, 64), (64, 64, 64, 128, 128), (64, 32, 32, 256, 256)][int(sys.argv[1]) if len(sys.argv) > 1 else 2]
secs = float(sys.argv[2]) if len(sys.argv) > 2 else 3.0
plan, keep = make(*cfg)
fl = 2.0 * cfg[0] * cfg[1] * cfg[2] * cfg[3] * 9 * cfg[4]
for _ in range(3):
    plan.run()
torch.cuda.synchronize()
th = threading.Thread(target=sampler); th.start()
time.sleep(0.3)
t_start = time.time()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
seg = []
while time.time() - t_start < secs:
    e0.record()
    for _ in range(200):
        plan.run()
    e1.record()
    torch.cuda.synchronize()
    seg.append((time.time() - t_start, e0.elapsed_time(e1) / 200))
t_end = time.time()
stop = True; th.join()
print("DBG", os.environ.get("PSSR_DBG"), "cfg", cfg)
for t, ms in seg[:3] + seg[len(seg)//2:len(seg)//2+2] + seg[-3:]:
    print(f"  t={t:5.2f}s  {ms*1000:7.1f} us  {fl/ms/1e9:6.0f} TF/s")
load = [(c, p) for (t, c, p) in samples if t_start + 0.5 < t < t_end]
idle = [(c, p) for (t, c, p) in samples if t < t_start - 0.05]
print("  idle samples", idle[:3])
if load:
    cs = sorted(c for c, _ in load); ps = sorted(p for _, p in load)
    print(f"  under load: n={len(load)} sm_mhz median {cs[len(cs)//2]} min {cs[0]} max {cs[-1]}; power median {ps[len(ps)//2]} max {ps[-1]}")
