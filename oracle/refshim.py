"""Import the *unmodified* reference (``/root/reference/pssr``) in the build container.

TEST INFRASTRUCTURE ONLY, usable where ``/root/reference`` exists (the build container) or where
``baseline/_ref`` holds the install of the reference's own wheel (``__graft_entry__.build()``).  The reference imports several third-party packages that
are not installed here (tifffile, czifile, scikit-image, timm, pytorch_msssim, skopt);
this module registers stand-ins in ``sys.modules`` -- real restatements for the routines
on the hot path (``oracle/thirdparty.py``), inert stubs for file I/O and training-only
packages -- and then imports ``pssr``.  It is used by ``tests/golden/gen_golden.py`` to
produce the committed golden vectors and by ``tests/test_oracle_vs_reference.py`` (skipped
when the reference is absent) to pin the oracle restatements.
"""
import os
import sys
import types

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _find_root():
    """The unmodified reference: $PSSR_REFERENCE_ROOT, the source checkout in the build container, or the copy installed from the
    reference's own wheel into baseline/_ref (git-ignored; it travels to the GPU box with the snapshot)."""
    for cand in (os.environ.get("PSSR_REFERENCE_ROOT"), "/root/reference", os.path.join(_REPO, "baseline", "_ref")):
        if cand and os.path.isdir(os.path.join(cand, "pssr")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _find_root()


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "pssr"))


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install_shims():
    from . import thirdparty as tp

    def _absent(name):
        try:
            __import__(name)
            return False
        except Exception:
            return True

    def _io_stub(*a, **k):
        raise RuntimeError("file I/O is bypassed in the oracle harness (in-memory datasets)")

    if _absent("tifffile"):
        _mod("tifffile", imread=_io_stub, imwrite=_io_stub)
    if _absent("czifile"):
        _mod("czifile", CziFile=_io_stub)
    if _absent("skimage"):
        sk = _mod("skimage")
        sk.metrics = _mod("skimage.metrics",
                          peak_signal_noise_ratio=tp.peak_signal_noise_ratio,
                          structural_similarity=tp.structural_similarity)
        sk.util = _mod("skimage.util", random_noise=tp.random_noise)
        sk.filters = _mod("skimage.filters", gaussian=tp.gaussian)
        sk.transform = _mod("skimage.transform", resize=tp.resize)
    if _absent("timm"):
        tm = _mod("timm")
        tm.layers = _mod("timm.layers", DropPath=tp.DropPath, LayerNorm2d=tp.LayerNorm2d,
                         EffectiveSEModule=tp.EffectiveSEModule,
                         to_2tuple=lambda x: (x, x) if not isinstance(x, (tuple, list)) else tuple(x),
                         trunc_normal_=lambda t, mean=0.0, std=1.0, a=-2.0, b=2.0: __import__('torch').nn.init.trunc_normal_(t, mean, std, a, b))      # timm's is torch's
        tm.models = _mod("timm.models", named_apply=tp.named_apply)
    if _absent("pytorch_msssim"):
        _mod("pytorch_msssim", SSIM=_io_stub, MS_SSIM=_io_stub, ssim=_io_stub, ms_ssim=_io_stub)
    if _absent("skopt"):
        sko = _mod("skopt", gp_minimize=_io_stub)
        sko.space = _mod("skopt.space", Real=_io_stub, Integer=_io_stub, Dimension=object)
        sko.utils = _mod("skopt.utils", use_named_args=_io_stub)


def import_reference():
    """Returns the reference ``pssr`` package (predict, data, crappifiers, util, models)."""
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    install_shims()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import pssr  # noqa: F401
    import pssr.crappifiers, pssr.data, pssr.util, pssr.predict, pssr.models  # noqa: F401,E401
    return pssr
