"""Developer probe: single-tap identity convs through the strip kernel, to see where pixels land."""
import sys, os
import torch
import torch.nn.functional as F
sys.path.insert(0, ".")
from pssr2_b200 import plan as P

def run(B, H, W, C, tap, prec="bf16", chan=False):
    plan = P.Plan(prec)
    dt = plan.tdtype
    x = torch.zeros(B, H, W, C, dtype=dt, device="cuda")
    # value encodes (n, y, x) in channel 0 ; channel c gets +c/64
    n_i = torch.arange(B, device="cuda").view(B, 1, 1, 1)
    y_i = torch.arange(H, device="cuda").view(1, H, 1, 1)
    x_i = torch.arange(W, device="cuda").view(1, 1, W, 1)
    c_i = torch.arange(C, device="cuda").view(1, 1, 1, C)
    x = (1 + n_i * 0 + y_i * 1.0 + x_i / 64.0 + 0 * c_i).to(dt).expand(B, H, W, C).contiguous()
    if chan:
        x = (1.0 + c_i + 0 * y_i + 0 * x_i + 0 * n_i).to(dt).expand(B, H, W, C).contiguous()
    w = torch.zeros(C, C, 3, 3, device="cuda")
    for c in range(C):
        w[c, c, tap // 3, tap % 3] = 1.0
    b = torch.zeros(C, device="cuda")
    wp = P.pack_weight([w], plan.dtype)
    out = torch.full((B, H, W, C), -7.0, dtype=dt, device="cuda")
    plan.conv([P.View(x)], [(0, 9, P.ceil_div(C, 64))], wp, b, P.View(out), Ho=H, Wo=W, B=B, act=P.ACT_NONE)
    plan.finalize()
    plan.run()
    torch.cuda.synchronize()
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w, b, padding=1).permute(0, 2, 3, 1)
    err = (out.float() - ref).abs()
    bad = err > 1e-2
    print(f"chan={chan} B{B} {H}x{W} C{C} tap{tap}: bad {int(bad.sum())}/{bad.numel()}  max {float(err.max()):.3f}")
    if bad.any():
        bp = bad.any(dim=3)[0]   # [H, W] map for image 0
        for yy in range(min(H, 20)):
            print("   ", "".join("X" if bp[yy, xx] else "." for xx in range(min(W, 64))))
        idx = bad.nonzero()[:6]
        for i in idx:
            n, yy, xx, c = [int(v) for v in i]
            print(f"    at n{n} y{yy} x{xx} c{c}: got {float(out[n,yy,xx,c]):.4f} want {float(ref[n,yy,xx,c]):.4f}")

for tap in (4, 3, 1, 0, 8):
    run(1, 16, 16, 64, tap, chan=True)
run(2, 8, 8, 64, 4, chan=True)
run(1, 32, 32, 128, 4, chan=True)
