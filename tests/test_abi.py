"""CPU tests: the C-ABI library loads, exports every symbol include/pssr_b200.h declares, and the ctypes
structures have exactly the C layout (checked by compiling the header with gcc).  No compute calls."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pssr_b200.h")


@pytest.fixture(scope="module")
def L():
    from pssr2_b200 import _lib
    _lib.build()
    return _lib


def test_library_exports_every_declared_symbol(L):
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    declared = set(re.findall(r"\b(pssr_[a-z0-9_]+)\s*\(", text))
    assert declared == set(L.SYMBOLS), declared ^ set(L.SYMBOLS)
    lib = L.lib()
    for name in declared:
        assert hasattr(lib, name), f"{name} not exported"
    assert lib.pssr_version().startswith(b"pssr_b200")
    assert lib.pssr_launch_count() == 0 or lib.pssr_launch_count() > 0


def test_ctypes_layout_matches_c(L, tmp_path):
    names = {"pssr_noise_stage_t": L.NoiseStage, "pssr_crappify_args_t": L.CrappifyArgs, "pssr_src_t": L.Src, "pssr_kseg_t": L.KSeg,
             "pssr_conv_desc_t": L.ConvDesc, "pssr_prep_desc_t": L.PrepDesc, "pssr_pool_desc_t": L.PoolDesc,
             "pssr_tail_desc_t": L.TailDesc, "pssr_tailsum_desc_t": L.TailSumDesc, "pssr_stem_desc_t": L.StemDesc, "pssr_ln_desc_t": L.LnDesc,
             "pssr_dwln_desc_t": L.DwLnDesc, "pssr_ese_desc_t": L.EseDesc, "pssr_cast8_desc_t": L.Cast8Desc, "pssr_resample_desc_t": L.ResampleDesc, "pssr_winattn_desc_t": L.WinAttnDesc, "pssr_op_t": L.Op}
    probes = [("pssr_crappify_args_t", "lr_frames"), ("pssr_crappify_args_t", "seed"), ("pssr_conv_desc_t", "weights"),
              ("pssr_conv_desc_t", "out_f32"), ("pssr_conv_desc_t", "tail_z"), ("pssr_conv_desc_t", "out_lo"), ("pssr_conv_desc_t", "weights8"), ("pssr_conv_desc_t", "resid_scale"),
              ("pssr_prep_desc_t", "im2col_lo"), ("pssr_cast8_desc_t", "out_choff"), ("pssr_kseg_t", "dilation"), ("pssr_resample_desc_t", "out_choff"), ("pssr_winattn_desc_t", "out_choff"), ("pssr_tailsum_desc_t", "out_u8"), ("pssr_stem_desc_t", "out_lo"), ("pssr_ln_desc_t", "out_lo"), ("pssr_dwln_desc_t", "out_lo"), ("pssr_ese_desc_t", "out_choff"), ("pssr_tail_desc_t", "out_u8"), ("pssr_noise_stage_t", "injected")]
    src = '#include <stdio.h>\n#include <stddef.h>\n#include "%s"\nint main(){\n' % HEADER
    for n in names:
        src += f'printf("{n} %zu\\n", sizeof({n}));\n'
    for n, f in probes:
        src += f'printf("{n}.{f} %zu\\n", offsetof({n}, {f}));\n'
    src += "return 0;}\n"
    c = tmp_path / "probe.c"
    c.write_text(src)
    exe = tmp_path / "probe"
    subprocess.run(["gcc", str(c), "-o", str(exe)], check=True)
    out = dict(line.split() for line in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for n, cls in names.items():
        assert int(out[n]) == ctypes.sizeof(cls), n
    for n, f in probes:
        assert int(out[f"{n}.{f}"]) == getattr(names[n], f).offset, (n, f)


def test_product_path_has_no_oracle_or_cpu_fallback():
    """The package must never import the oracle, and must fail loudly without CUDA."""
    import torch
    pkg = os.path.join(ROOT, "pssr2_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            assert "oracle" not in open(os.path.join(pkg, f)).read().replace("oracle/", ""), f
    from pssr2_b200.models import ResUNet
    with pytest.raises(RuntimeError):
        ResUNet(hidden=[64, 128]).eval()(torch.zeros(1, 1, 32, 32))
    from pssr2_b200.predict import predict_images

    class DS:
        is_lr, val_idx = False, [0]
    with pytest.raises(RuntimeError):
        predict_images(ResUNet(hidden=[64, 128]), DS(), device="cpu", out_dir=None)
