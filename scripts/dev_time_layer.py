"""Developer timing of single conv layers through the plan API (not a bench)."""
import sys, os
import torch
sys.path.insert(0, ".")
from pssr2_b200 import plan as P

def time_layer(B, H, W, Cin, Cout, prec="fp16", reps=10):
    plan = P.Plan(prec)
    dt = plan.tdtype
    x = torch.randn(B, H, W, Cin, device="cuda").to(dt)
    w = torch.randn(Cout, Cin, 3, 3, device="cuda") / (3 * Cin ** 0.5)
    b = torch.zeros(Cout, device="cuda")
    wp = P.pack_weight([w], plan.dtype)
    out = torch.zeros(B, H, W, Cout, dtype=dt, device="cuda")
    plan.conv([P.View(x)], [(0, 9, P.ceil_div(Cin, 64))], wp, b, P.View(out), Ho=H, Wo=W, B=B, act=P.ACT_RELU)
    plan.finalize()
    for _ in range(3):
        plan.run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        plan.run()
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / reps
    fl = 2.0 * B * H * W * Cin * 9 * Cout
    print(f"  {B}x{H}x{W} {Cin}->{Cout}: {t*1000:8.1f} us  {fl/t/1e9:7.0f} TF/s", flush=True)

print("PSSR_DBG =", os.environ.get("PSSR_DBG"), "T =", os.environ.get("PSSR_STRIP_T"), "V1 =", os.environ.get("PSSR_CONV_V1"))
CFGS = [(64, 128, 128, 64, 64), (64, 64, 64, 128, 128), (64, 32, 32, 256, 256), (64, 16, 16, 512, 512), (64, 8, 8, 1024, 1024)]
if len(sys.argv) > 1:
    CFGS = [CFGS[int(a)] for a in sys.argv[1:]]
for cfg in CFGS:
    time_layer(*cfg)
