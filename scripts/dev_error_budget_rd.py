"""CPU study: which 16-bit quantisation sites of the RDResUNet plan dominate the max-abs error vs fp32 (dev tool, not shipped).
Sites already compensated by the fp16c plan (input, Reconstruction, last respass weights) can be switched off to see what is left."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.nn.functional as F
from oracle import models as OM
from pssr2_b200.models import RDResUNet
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gpu_net import _randomise_bn, _randomise_rd

DW_HI_ONLY = False
ON = None        # None: every site rounds; else only sites matching
OFF = set()
SPLIT = set()


def _match(site, pats):
    return any(site == p or (p.endswith("*") and site.startswith(p[:-1])) for p in pats)


def q(t, site):
    if _match(site, OFF) or (ON is not None and not _match(site, ON)):
        return t
    hi = t.half().float()
    if _match(site, SPLIT):
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


def rd_block(sd, p, x, tag):
    L = p + ".layers.layers"
    c = x.shape[1]
    xin = x.half().float() if (DW_HI_ONLY and tag.startswith("s0")) else x
    h = F.conv2d(xin, sd[L + ".0.weight"], sd[L + ".0.bias"], padding=3, groups=c)          # fp32 weights on CUDA cores
    h = q(OM._ln2d(h, sd[L + ".1.weight"], sd[L + ".1.bias"]), f"a.{tag}.dwln")
    h = F.conv2d(h, q(sd[L + ".2.weight"], f"w.{tag}.expand"), sd[L + ".2.bias"])
    h = q(F.gelu(h), f"a.{tag}.mid")
    h = F.conv2d(h, q(sd[L + ".4.weight"], f"w.{tag}.project"), sd[L + ".4.bias"])
    if L + ".5.fc.weight" in sd:
        h = q(h, f"a.{tag}.g")
        se = h.mean((2, 3), keepdim=True)
        se = F.conv2d(se, sd[L + ".5.fc.weight"], sd[L + ".5.fc.bias"])
        h = h * (F.relu6(se + 3.0) / 6.0)
    if p + ".gamma" in sd:
        h = h * sd[p + ".gamma"].view(1, -1, 1, 1)
    return q(h, f"a.{tag}.out")


def rdnet(sd, x, ds_blocks, P="encoder"):
    w = sd[P + ".stem.stem.0.weight"]
    x = F.conv2d(x, w, sd[P + ".stem.stem.0.bias"], stride=w.shape[-1])
    x = q(OM._ln2d(x, sd[P + ".stem.stem.1.weight"], sd[P + ".stem.stem.1.bias"]), "a.stem")
    skips = []
    for i, ds in enumerate(ds_blocks):
        if ds:
            skips.append(x)
        S = f"{P}.dense_stages.{i}"
        k = 0
        if f"{S}.0.weight" in sd:
            x = q(OM._ln2d(x, sd[f"{S}.0.weight"], sd[f"{S}.0.bias"]), f"a.s{i}.tln")
            wt = sd[f"{S}.1.weight"]
            x = q(F.conv2d(x, q(wt, f"w.s{i}.trans"), sd[f"{S}.1.bias"], stride=wt.shape[-1]), f"a.s{i}.trans")
            k = 2
        feats = [x]
        j = 0
        while f"{S}.{k}.dense_block{j}.layers.layers.0.weight" in sd:
            feats.append(rd_block(sd, f"{S}.{k}.dense_block{j}", torch.cat(feats, 1), f"s{i}.b{j}"))
            j += 1
        x = torch.cat(feats, 1)
    return skips + [x]


def forward(sd, x, ds_blocks=(False, True, True, False, False, False, True), patch=2):
    sd = {k: v.float() for k, v in sd.items() if v.is_floating_point()}
    xn = OM._input_norm(sd, x.float())
    skips = [q(xn, "a.input")] + rdnet(sd, xn, ds_blocks)            # the stem kernel reads the fp32 input itself
    n_dec = 4
    ratios = [1] + [2] * (n_dec - 1) + [patch]
    for i in range(n_dec):
        x = torch.cat([x, skips.pop()], 1) if i != 0 else skips.pop()
        x = resblock(sd, f"decoder.{i}", x)
        x = F.pixel_shuffle(x, ratios[i + 1])
    x = torch.cat([x, skips.pop()], 1)
    h = F.relu(F.conv2d(x, q(sd["reconstruction.pre.weight"], "w.recon.pre"), sd["reconstruction.pre.bias"], padding=1))
    h = F.pixel_shuffle(q(h, "a.recon.pre"), 4)
    y = F.conv2d(h, q(sd["reconstruction.conv.weight"], "w.recon.conv"), sd["reconstruction.conv.bias"], padding=1)
    return y * 128 + 128


COMP = ["a.input", "w.recon.pre", "a.recon.pre", "w.recon.conv", "w.decoder.3.res"]

if __name__ == "__main__":
    torch.manual_seed(0)
    model = RDResUNet().eval(); _randomise_bn(model)
    if "--default-gamma" not in sys.argv:
        _randomise_rd(model)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    rng = np.random.default_rng(0)
    x = torch.tensor(rng.integers(0, 256, (1, 1, 128, 128)).astype(np.float32))
    want = OM.rdresunet_forward(sd, x)

    def run(label, on=None, off=(), split=()):
        global ON
        ON = on; OFF.clear(); OFF.update(off); SPLIT.clear(); SPLIT.update(split)
        d = (forward(sd, x) - want).abs()
        print(f"{label:50s} max-abs {float(d.max()):.5f}  var*1e6 {float((d**2).mean())*1e6:8.3f}", flush=True)

    run("all fp16")
    run("fp16c today (compensated sites exact)", off=COMP)
    groups = ["a.stem", "a.s0.*", "w.s0.*", "a.s1.*", "w.s1.*", "a.s2.*", "w.s2.*", "a.s3.*", "a.s4.*", "a.s5.*", "a.s6.*",
              "w.s3.*", "w.s4.*", "w.s5.*", "w.s6.*", "a.decoder.0*", "a.decoder.1*", "a.decoder.2*", "w.decoder.0*", "w.decoder.1*",
              "w.decoder.2*", "a.decoder.3.0", "a.decoder.3.1", "a.decoder.3.2", "a.decoder.3.3", "w.decoder.3.0", "w.decoder.3.1",
              "w.decoder.3.2", "w.decoder.3.3"]
    for g in (groups if "--sites" in sys.argv else []):
        run("only " + g, on=[g])
    for j in (range(3) if "--sites" in sys.argv else []):
        for s in ("dwln", "mid", "out"):
            run(f"only a.s0.b{j}.{s}", on=[f"a.s0.b{j}.{s}"])
        for s in ("expand", "project"):
            run(f"only w.s0.b{j}.{s}", on=[f"w.s0.b{j}.{s}"])
    S0 = ["a.stem", "a.s0.*", "w.s0.*"]
    run("fp16c + stage 0 split", off=COMP, split=S0)
    run("fp16c + stage 0 + final split", off=COMP, split=S0 + ["a.decoder.3.3"])
    run("fp16c + stage 0 + final + s1 split", off=COMP, split=S0 + ["a.decoder.3.3", "a.s1.*", "w.s1.*"])
    run("fp16c + stage 0 acts only + final split", off=COMP, split=["a.stem", "a.s0.*", "a.decoder.3.3"])
    DW_HI_ONLY = True
    run("fp16c + stage 0 + final split, dw reads hi only", off=COMP, split=S0 + ["a.decoder.3.3"])
    run("  ... and expand/project weights single (acts split)", off=COMP, split=["a.stem", "a.s0.*", "a.decoder.3.3"])
