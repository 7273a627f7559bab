import sys
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

