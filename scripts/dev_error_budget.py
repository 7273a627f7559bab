"""CPU study: which 16-bit quantisation sites of the ResUNet plan dominate the max-abs error vs fp32
(uses the oracle's emulation structure with per-site switches; dev tool, not shipped)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.nn.functional as F
from oracle import models as OM
from pssr2_b200.models import ResUNet
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gpu_net import _randomise_bn

OFF = set()      # site names whose quantisation is disabled
SPLIT = set()    # site names quantised as hi+lo (two fp16 terms)

def q(t, site):
    if site in OFF or any(site.startswith(p[:-1]) for p in OFF if p.endswith("*")):
        return t
    hi = t.half().float()
    if site in SPLIT or any(site.startswith(p[:-1]) for p in SPLIT if p.endswith("*")):
        return hi + (t - hi).half().float()
    return hi

def resblock(sd, prefix, x):
    n = OM._n_convs(sd, prefix)
    h = x
    for i in range(n):
        w = sd[f"{prefix}.conv.{3*i}.weight"]; b = sd[f"{prefix}.conv.{3*i}.bias"]
        s, t = OM._bn_fold(sd, f"{prefix}.conv.{3*i+1}")
        wf = q(w * s.view(-1, 1, 1, 1), f"w.{prefix}.{i}")
        acc = F.conv2d(h, wf, None, padding=1) + (b * s + t).view(1, -1, 1, 1)
        if i + 1 < n:
            h = q(F.relu(acc), f"a.{prefix}.{i}")
        else:
            wr = q(sd[f"{prefix}.respass.weight"], f"w.{prefix}.res")
            acc = acc + F.conv2d(x, wr, sd[f"{prefix}.respass.bias"])
            h = q(F.relu(acc), f"a.{prefix}.{i}")
    return h

def forward(sd, x):
    sd = {k: v.float() for k, v in sd.items() if v.is_floating_point()}
    x = q(OM._input_norm(sd, x.float()), "a.input")
    skips = [x]
    for i in range(5):
        x = resblock(sd, f"encoder.{i}", x)
        if i < 4:
            skips.append(x); x = F.max_pool2d(x, 2)
    for i in range(4):
        x = torch.cat([F.pixel_shuffle(x, 2), skips.pop()], 1)
        x = resblock(sd, f"decoder.{i}", x)
    x = torch.cat([x, skips.pop()], 1)
    h = F.relu(F.conv2d(x, q(sd["reconstruction.pre.weight"], "w.recon.pre"), sd["reconstruction.pre.bias"], padding=1))
    h = F.pixel_shuffle(q(h, "a.recon.pre"), 4)
    y = F.conv2d(h, q(sd["reconstruction.conv.weight"], "w.recon.conv"), sd["reconstruction.conv.bias"], padding=1)
    return y * 128 + 128

if __name__ == "__main__":
    torch.manual_seed(0)
    model = ResUNet().eval(); _randomise_bn(model)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    rng = np.random.default_rng(0)
    x = torch.tensor(rng.integers(0, 256, (2, 1, 128, 128)).astype(np.float32))
    want = OM.resunet_forward(sd, x)
    def run(label, off=(), split=()):
        OFF.clear(); OFF.update(off); SPLIT.clear(); SPLIT.update(split)
        d = (forward(sd, x) - want).abs()
        print(f"{label:60s} max-abs {float(d.max()):.5f}  rms {float((d**2).mean().sqrt()):.5f}")
    run("all fp16")
    run("weights exact", off=["w.*"])
    run("activations exact", off=["a.*"])
    run("recon (w+a) exact", off=["w.recon*", "a.recon*"])
    run("recon.pre act exact", off=["a.recon.pre"])
    run("recon.conv w exact", off=["w.recon.conv"])
    run("recon.pre w exact", off=["w.recon.pre"])
    run("decoder.3 + recon exact", off=["w.recon*", "a.recon*", "w.decoder.3*", "a.decoder.3*"])
    run("input exact", off=["a.input"])
    run("encoder exact", off=["w.encoder*", "a.encoder*", "a.input"])
    run("decoder exact", off=["w.decoder*", "a.decoder*"])
    run("only recon quantised", off=["w.encoder*", "a.encoder*", "a.input", "w.decoder*", "a.decoder*"])
    run("split recon.pre w + a", split=["w.recon.pre", "a.recon.pre"])
    run("split recon.pre w + a, dec3 last act", split=["w.recon.pre", "a.recon.pre", "a.decoder.3.3", "a.input"])
