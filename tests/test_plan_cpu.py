"""CPU-only checks of the host logic: the plan a module emits, interpreted with PyTorch fp32 ops on the
packed weights (tests/plan_interp.py), must reproduce the oracle forward -- i.e. packing, BatchNorm
folding, K scheduling, concat offsets and pixel-shuffle permutation are right before any GPU runs."""
import numpy as np
import pytest
import torch

from oracle.models import resunet_forward
from pssr2_b200 import plan as P
from tests.plan_interp import run_records
from tests.test_gpu_net import _randomise_bn


@pytest.fixture
def dry_run():
    P.DRY_RUN = True
    yield
    P.DRY_RUN = False


@pytest.mark.parametrize("cfg", [dict(), dict(channels=[3, 3], hidden=[64, 128, 256], scale=2, depth=1),
                                 dict(channels=[5, 1], hidden=[64, 128], scale=8, depth=0),
                                 dict(channels=[9, 1], hidden=[64, 128], scale=2, depth=1)])   # > 7 channels: no im2col, 3x3 over the input
@pytest.mark.parametrize("prec", ["fp16", "fp16c"])
def test_resunet_plan_matches_oracle_on_cpu(dry_run, cfg, prec):
    from pssr2_b200.models import ResUNet
    torch.manual_seed(0)
    model = ResUNet(**cfg).eval()
    _randomise_bn(model)
    cin = model.channels[0]
    small = bool(cfg)
    B, H, W = (2, 32, 48) if small else (1, 32, 32)
    x = torch.tensor(np.random.default_rng(0).integers(0, 256, (B, cin, H, W)).astype(np.float32))
    want = resunet_forward(model.state_dict(), x)
    model.precision = prec
    st = model._build(x.shape, x.dtype, torch.device("cpu"))
    st["x"].copy_(x)
    run_records(st["plan"])
    got = st["out"]
    assert got.shape == want.shape
    # 16-bit operand rounding only (fp16: ~1e-2 on the 0..255 scale); an indexing / packing bug gives O(1..100).
    # The compensated plan (hi + lo on the shallow full-resolution path) must land below the 1e-2 bar of the north star.
    err = float((got - want).abs().max())
    print(f"{cfg} {prec}: max-abs {err:.5f}")
    # (the reduced test nets put relatively more weight on the one uncompensated term left, the last decoder output feeding
    # Reconstruction.pre; the default-depth models the 1e-2 bar is stated for are asserted in tests/test_gpu_net.py)
    assert err < ((1e-2 if not cfg else 2e-2) if prec == "fp16c" else 3e-2)
    c = got.shape[1] // 2
    assert torch.equal(st["out_u8"], got[:, c:c + 1].clamp(0, 255).to(torch.uint8))


def test_resunet_plan_window48_tail_on_cpu(dry_run):
    """W % 128 == 0 and scale 4: the fused tail uses the PSSR_TAIL_WINDOW48 layout (include/pssr_b200.h); its CPU statement in
    tests/plan_interp.py must reproduce the oracle forward as well."""
    from pssr2_b200.models import ResUNet
    torch.manual_seed(3)
    model = ResUNet(hidden=[64, 128], depth=0).eval()
    _randomise_bn(model)
    x = torch.tensor(np.random.default_rng(3).integers(0, 256, (2, 1, 4, 128)).astype(np.float32))
    want = resunet_forward(model.state_dict(), x)
    for prec, tol in (("fp16", 3e-2), ("fp16c", 2e-2)):      # two levels, depth 0: a packing test, not a precision claim
        model.precision = prec
        st = model._build(x.shape, x.dtype, torch.device("cpu"))
        assert any(k == "tailsum" and r["layout"] == 1 for k, r in st["plan"].records)
        st["x"].copy_(x)
        run_records(st["plan"])
        assert float((st["out"] - want).abs().max()) < tol


@pytest.mark.parametrize("cfg", [dict(), dict(hidden=[128, 64], growth_rates=[32, 40, 64], ds_blocks=[False, True, False],
                                                ese_blocks=[False, True, True], n_blocks=[2, 1, 2], rdnet_init=64, scale=2, depth=1)])
def test_rdresunet_plan_matches_oracle_on_cpu(dry_run, cfg):
    from oracle.models import rdresunet_forward
    from pssr2_b200.models import RDResUNet
    torch.manual_seed(0)
    model = RDResUNet(**cfg).eval()
    _randomise_bn(model)
    with torch.no_grad():   # make layer-scale / LayerNorm parameters non-trivial (default gamma = 1e-6 hides the dense blocks)
        g = torch.Generator().manual_seed(5)
        for n, p in model.named_parameters():
            if n.endswith("gamma"):
                p.copy_(torch.rand(p.shape, generator=g) * 0.5 + 0.25)
            elif "encoder" in n and p.dim() == 1 and ("weight" in n) and p.shape[0] > 1 and "layers.0" not in n and "fc" not in n:
                p.mul_(torch.rand(p.shape, generator=g) * 0.4 + 0.8)
    B, H, W = (1, 64, 64) if not cfg else (2, 32, 48)
    x = torch.tensor(np.random.default_rng(0).integers(0, 256, (B, 1, H, W)).astype(np.float32))
    ds = cfg.get("ds_blocks", (False, True, True, False, False, False, True))
    want = rdresunet_forward(model.state_dict(), x, ds_blocks=ds)
    model.precision = "fp16"
    st = model._build(x.shape, x.dtype, torch.device("cpu"))
    st["x"].copy_(x)
    run_records(st["plan"])
    got = st["out"]
    assert got.shape == want.shape
    err = float((got - want).abs().max())
    assert err < 6e-2, err      # 16-bit activations through ~60 layers incl. LayerNorm/GELU; packing bugs give O(1..100)


def test_rdresunet_compensated_plan_on_cpu(dry_run):
    """precision "fp16c": RDNet stage 0 (stem, dwconv + LayerNorm, expand / GELU / project) as (hi, lo) pairs, the low halves of the
    stage-0 skip in the last respass and `final_lo` as a second e5m2 term of Reconstruction.pre -- the emitted plan must land below
    the north star's 1e-2 where the single-pass plan sits at ~3e-2 (scripts/dev_error_budget_rd.py)."""
    from oracle.models import rdresunet_forward
    from pssr2_b200.models import RDResUNet
    from tests.test_gpu_net import _randomise_rd
    torch.manual_seed(0)
    model = RDResUNet().eval()
    _randomise_bn(model)
    _randomise_rd(model)
    x = torch.tensor(np.random.default_rng(0).integers(0, 256, (1, 1, 128, 128)).astype(np.float32))
    want = rdresunet_forward(model.state_dict(), x)
    model.precision = "fp16c"
    st = model._build(x.shape, x.dtype, torch.device("cpu"))
    recs = st["plan"].records
    assert any(k == "stem" and r["out_lo"] is not None for k, r in recs) and any(k == "dwln" and r["src_lo"] is not None for k, r in recs)
    assert any(k == "conv" and any(len(sg) > 3 and sg[3] == 1 and sg[2] == 2 for sg in r["segs"]) for k, r in recs)     # [final8 | final_lo8]
    st["x"].copy_(x)
    run_records(st["plan"])
    err = float((st["out"] - want).abs().max())
    assert err < 1e-2, err


def test_variant_plans_match_reference_golden_on_cpu(dry_run):
    """Atrous residual blocks (dilated taps, pre-activation BatchNorm -> ReLU, chained partial sums for more than three branches)
    and PSP pooling (k x k max pool, bilinear enlargement, block-diagonal 1x1) as emitted plans, interpreted on the CPU, against the
    reference's own outputs (tests/golden/net_variants.npz)."""
    import os
    from tests.test_oracle import G, VARIANT_CASES, variant_model, variant_oracle
    g = np.load(os.path.join(G, "net_variants.npz"))
    for tag, cls, kw in VARIANT_CASES:
        m = variant_model(tag, cls, kw, g)
        x = torch.as_tensor(g[f"{tag}_x"])
        want = torch.as_tensor(g[f"{tag}_y"]) if m is not None else None
        if m is None:        # seeded init not reproducible here: fall back to the oracle on fresh weights
            import pssr2_b200.models as M
            m = getattr(M, cls)(**kw).eval()
            _randomise_bn(m)
            want = variant_oracle(cls, kw, m.state_dict(), x)
        st = m._build(x.shape, x.dtype, torch.device("cpu"))
        kinds = {k for k, _ in st["plan"].records}
        assert "resample" in kinds
        if kw.get("dilations"):
            assert any(k == "conv" and any(len(sg) > 4 and sg[4] > 1 for sg in r["segs"]) for k, r in st["plan"].records)
        st["x"].copy_(x)
        run_records(st["plan"])
        err = float((st["out"] - want).abs().max())
        assert err < 3e-2, (tag, err)      # single-pass fp16 plan (the variants carry no compensation terms)


def test_swinir_plan_matches_reference_golden_on_cpu(dry_run):
    """SwinIR as an emitted plan (1x1 GEMMs for the linear layers with residual / GELU epilogues, LayerNorm, shifted-window attention,
    pixel-shuffle upsampling, CUDA-core last convolution), interpreted on the CPU, against the reference's own outputs."""
    import os
    from oracle.models import swinir_forward
    from tests.test_oracle import G, SWINIR_CASES, swinir_model
    g = np.load(os.path.join(G, "swinir.npz"))
    for tag, kw in SWINIR_CASES:
        m = swinir_model(tag, kw, g)
        x = torch.as_tensor(g[f"{tag}_x"])
        want = torch.as_tensor(g[f"{tag}_y"]) if m is not None else None
        if m is None:
            from pssr2_b200.models import SwinIR
            m = SwinIR(**kw).eval()
            want = swinir_forward(m.state_dict(), x)
        st = m._build(x.shape, x.dtype, torch.device("cpu"))
        assert sum(k == "winattn" for k, _ in st["plan"].records) == sum(kw["depths"])
        st["x"].copy_(x)
        run_records(st["plan"])
        err = float((st["out"] - want).abs().max())
        assert err < 8e-2, (tag, err)       # single-pass fp16 over 0..255-scale features (conv_first output is not normalised)
        c = st["out"].shape[1] // 2
        assert torch.equal(st["out_u8"], st["out"][:, c:c + 1].clamp(0, 255).to(torch.uint8))


def test_weights_stamp_sees_every_kind_of_update():
    """The plan cache key of a model (`_PlanModule._weights_stamp`, taken on every forward without walking the module tree): in-place
    parameter / buffer updates, replaced Parameter objects and storage moves all change it; copies and pickles of a model carry the
    parameters but never its device plans."""
    import copy
    import pickle
    from pssr2_b200.models import RDResUNet, ResUNet
    for M in (ResUNet, RDResUNet):
        m = M().eval()
        s0 = m._weights_stamp()
        assert m._weights_stamp() == s0 and "_stamp_modules" in m.__dict__
        with torch.no_grad():
            next(iter(m.parameters())).mul_(1.0)                        # version bump, same values
        s1 = m._weights_stamp()
        assert s1 != s0
        bufs = [b for b in m.buffers() if b.is_floating_point()]
        bufs[-1].add_(1.0)
        s2 = m._weights_stamp()
        assert s2 != s1
        m.reconstruction.conv.weight = torch.nn.Parameter(m.reconstruction.conv.weight.detach().clone())       # a new object
        s3 = m._weights_stamp()
        assert s3 != s2
        m.invalidate()
        assert "_stamp_modules" not in m.__dict__ and m._weights_stamp() == s3
    m = ResUNet().eval()
    m._weights_stamp()
    m._plans[("fake",)] = {"plan": object()}
    for c in (copy.deepcopy(m), pickle.loads(pickle.dumps(m))):
        assert c._plans == {} and "_stamp_modules" not in c.__dict__
        assert all(torch.equal(a, b) for a, b in zip(m.state_dict().values(), c.state_dict().values()))
    assert ("fake",) in m._plans
    del m._plans[("fake",)]
