"""ctypes binding of ``libpssr_b200.so`` (C ABI declared in ``include/pssr_b200.h``).

There is deliberately NO fallback: if the shared library is missing or a call fails, a
``RuntimeError`` is raised.  ``build()`` compiles the library in-tree with nvcc for sm_100a.
"""
import ctypes
import os
import subprocess
import sys
from ctypes import (POINTER, Structure, Union, c_char_p, c_double, c_float, c_int32, c_int64, c_uint8,
                    c_uint64, c_void_p)

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_HERE, "libpssr_b200.so")
SOURCES = ["api.cu", "conv_igemm.cu", "conv_v3.cu", "tiff_io.cu", "net_aux.cu", "rdnet.cu", "swin.cu", "crappify.cu", "stitch.cu", "metrics.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--use_fast_math=false"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a and link ``libpssr_b200.so`` next to this file."""
    srcs = [os.path.join(_CSRC, s) for s in SOURCES]
    deps = srcs + [os.path.join(_CSRC, h) for h in ("common.cuh", "plan.h")] + [
        os.path.join(_HERE, "..", "include", "pssr_b200.h")]
    if not force and os.path.exists(LIB_PATH):
        newest = max(os.path.getmtime(p) for p in deps)
        if os.path.getmtime(LIB_PATH) >= newest:
            return LIB_PATH
    objdir = os.path.join(_HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    flags = [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")] + os.environ.get("PSSR_NVCC_EXTRA", "").split()   # developer builds
    procs = []
    objs = []
    hdr_time = max(os.path.getmtime(p) for p in deps[len(srcs):])
    for s in srcs:
        o = os.path.join(objdir, os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        if not force and not os.environ.get("PSSR_NVCC_EXTRA") and os.path.exists(o) and os.path.getmtime(o) >= max(os.path.getmtime(s), hdr_time):
            continue                      # this object is newer than its source and every header
        cmd = [_nvcc()] + flags + ["-c", s, "-o", o]
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("nvcc failed: %s\n%s" % (" ".join(cmd), out.decode(errors="replace")))
    tmp = os.path.join(objdir, "libpssr_b200.so.tmp")      # linked aside and renamed: a reader never sees a half-written library
    cmd = [_nvcc(), "-shared", "-o", tmp] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    if r.returncode != 0:
        raise RuntimeError("link failed: %s\n%s" % (" ".join(cmd), r.stdout.decode(errors="replace")))
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


# ------------------------------------------------------------------------------ structs
class NoiseStage(Structure):
    _fields_ = [("kind", c_int32), ("rng", c_int32), ("intensity", c_double), ("gain", c_double),
                ("mix_in_f32", c_int32), ("reserved", c_int32), ("injected", c_void_p)]


class CrappifyArgs(Structure):
    _fields_ = [("sheets", c_void_p), ("n_sheets", c_int32), ("elem_bytes", c_int32),
                ("sheet_h", c_int32), ("sheet_w", c_int32),
                ("tile_sheet", c_void_p), ("tile_frame", c_void_p), ("tile_y", c_void_p),
                ("tile_x", c_void_p), ("tile_vh", c_void_p), ("tile_vw", c_void_p),
                ("n_tiles", c_int32), ("frames", c_int32), ("hr_res", c_int32), ("lr_scale", c_int32),
                ("stages", NoiseStage * 4), ("n_stages", c_int32), ("clip_between", c_int32),
                ("seed", c_uint64), ("tile_index0", c_uint64),
                ("lr_out", c_void_p), ("hr_out", c_void_p), ("hr_u8_out", c_void_p),
                ("hr_frame0", c_int32), ("hr_frames", c_int32), ("lr_frame0", c_int32), ("lr_frames", c_int32),
                ("sheet_hs", c_void_p), ("sheet_ws", c_void_p), ("tile_xf", c_void_p)]


class Src(Structure):
    _fields_ = [("base", c_void_p), ("channels", c_int32), ("cstride", c_int32), ("H", c_int32),
                ("W", c_int32), ("B", c_int32), ("reserved", c_int32)]


class KSeg(Structure):
    _fields_ = [("src", c_int32), ("taps", c_int32), ("cblocks", c_int32), ("fmt", c_int32), ("dilation", c_int32)]


class ConvDesc(Structure):
    _fields_ = [("srcs", Src * 4), ("n_srcs", c_int32), ("segs", KSeg * 6), ("n_segs", c_int32),
                ("weights", c_void_p), ("bias", c_void_p), ("n", c_int32), ("n_valid", c_int32),
                ("Ho", c_int32), ("Wo", c_int32), ("B", c_int32), ("out", c_void_p),
                ("out_cstride", c_int32), ("out_choff", c_int32), ("shuffle", c_int32), ("act", c_int32),
                ("out_scale", c_void_p), ("out_f32", c_void_p), ("tail_weight", c_void_p), ("tail_z", c_void_p),
                ("tail_layout", c_int32), ("tail_flags", c_int32), ("out_lo", c_void_p),
                ("out_lo_cstride", c_int32), ("out_lo_choff", c_int32), ("weights8", c_void_p),
                ("resid", c_void_p), ("resid_cstride", c_int32), ("resid_choff", c_int32), ("resid_scale", c_float), ("reserved3", c_int32)]


class Cast8Desc(Structure):
    _fields_ = [("in_", c_void_p), ("in_cstride", c_int32), ("in_choff", c_int32), ("C", c_int32), ("B", c_int32), ("H", c_int32),
                ("W", c_int32), ("scale", c_float), ("out", c_void_p), ("out_cstride", c_int32), ("out_choff", c_int32)]


class PrepDesc(Structure):
    _fields_ = [("x", c_void_p), ("x_u8", c_int32), ("B", c_int32), ("C", c_int32), ("H", c_int32),
                ("W", c_int32), ("scale", c_void_p), ("shift", c_void_p), ("im2col", c_void_p), ("im2col_lo", c_void_p),
                ("xnorm_f32", c_void_p), ("cols", c_int32), ("centre_only", c_int32)]


class PoolDesc(Structure):
    _fields_ = [("in_", c_void_p), ("in_cstride", c_int32), ("in_choff", c_int32), ("out", c_void_p),
                ("out_cstride", c_int32), ("out_choff", c_int32), ("B", c_int32), ("H", c_int32),
                ("W", c_int32), ("C", c_int32)]


class TailDesc(Structure):
    _fields_ = [("in_", c_void_p), ("cstride", c_int32), ("C", c_int32), ("B", c_int32), ("H", c_int32),
                ("W", c_int32), ("weight", c_void_p), ("bias", c_void_p), ("Cout", c_int32),
                ("mul", c_float), ("add", c_float), ("out_f32", c_void_p), ("out_u8", c_void_p)]


class TailSumDesc(Structure):
    _fields_ = [("z", c_void_p), ("B", c_int32), ("H", c_int32), ("W", c_int32), ("r", c_int32), ("bias", c_float),
                ("mul", c_float), ("add", c_float), ("layout", c_int32), ("out_f32", c_void_p), ("out_u8", c_void_p)]


class StemDesc(Structure):
    _fields_ = [("x", c_void_p), ("x_u8", c_int32), ("B", c_int32), ("C", c_int32), ("H", c_int32), ("W", c_int32),
                ("in_scale", c_void_p), ("in_shift", c_void_p), ("patch", c_int32), ("Cout", c_int32), ("weight", c_void_p),
                ("bias", c_void_p), ("ln_w", c_void_p), ("ln_b", c_void_p), ("eps", c_float), ("reserved", c_int32),
                ("out", c_void_p), ("out_cstride", c_int32), ("out_choff", c_int32), ("out_lo", c_void_p)]


class LnDesc(Structure):
    _fields_ = [("in_", c_void_p), ("in_cstride", c_int32), ("in_choff", c_int32), ("C", c_int32), ("B", c_int32), ("H", c_int32),
                ("W", c_int32), ("s2d", c_int32), ("w", c_void_p), ("b", c_void_p), ("eps", c_float), ("reserved", c_int32),
                ("out", c_void_p), ("out_cstride", c_int32), ("out_choff", c_int32), ("in_lo", c_void_p), ("out_lo", c_void_p)]


class DwLnDesc(Structure):
    _fields_ = [("in_", c_void_p), ("in_cstride", c_int32), ("in_choff", c_int32), ("C", c_int32), ("B", c_int32), ("H", c_int32),
                ("W", c_int32), ("reserved", c_int32), ("dw_w", c_void_p), ("dw_b", c_void_p), ("ln_w", c_void_p), ("ln_b", c_void_p),
                ("eps", c_float), ("reserved2", c_int32), ("out", c_void_p), ("out_cstride", c_int32), ("out_choff", c_int32),
                ("in_lo", c_void_p), ("out_lo", c_void_p)]


class ResampleDesc(Structure):
    _fields_ = [("in_", c_void_p), ("in_cstride", c_int32), ("in_choff", c_int32), ("C", c_int32), ("B", c_int32), ("H", c_int32),
                ("W", c_int32), ("mode", c_int32), ("k", c_int32), ("Ho", c_int32), ("Wo", c_int32), ("relu", c_int32),
                ("scale", c_void_p), ("shift", c_void_p), ("out", c_void_p), ("out_cstride", c_int32), ("out_choff", c_int32)]


class WinAttnDesc(Structure):
    _fields_ = [("qkv", c_void_p), ("cstride", c_int32), ("C", c_int32), ("heads", c_int32), ("B", c_int32), ("H", c_int32), ("W", c_int32),
                ("ws", c_int32), ("shift", c_int32), ("scale", c_float), ("reserved", c_int32), ("biasT", c_void_p), ("out", c_void_p),
                ("out_cstride", c_int32), ("out_choff", c_int32)]


class EseDesc(Structure):
    _fields_ = [("in_", c_void_p), ("in_cstride", c_int32), ("C", c_int32), ("B", c_int32), ("H", c_int32), ("W", c_int32),
                ("reserved", c_int32), ("fc_w", c_void_p), ("fc_b", c_void_p), ("gamma", c_void_p), ("gate_ws", c_void_p),
                ("out", c_void_p), ("out_cstride", c_int32), ("out_choff", c_int32)]


class _OpU(Union):
    _fields_ = [("conv", ConvDesc), ("prep", PrepDesc), ("pool", PoolDesc), ("tail", TailDesc), ("tailsum", TailSumDesc), ("stem", StemDesc), ("ln", LnDesc), ("dwln", DwLnDesc), ("ese", EseDesc), ("cast8", Cast8Desc), ("resample", ResampleDesc), ("winattn", WinAttnDesc),
                ("pad", c_uint8 * 512)]


class Op(Structure):
    _fields_ = [("kind", c_int32), ("reserved", c_int32), ("u", _OpU)]


OP_CONV, OP_PREP, OP_MAXPOOL, OP_TAIL, OP_TAILSUM = 1, 2, 3, 4, 9
OP_DWCONV_LN, OP_LAYERNORM, OP_ESE, OP_STEM, OP_CAST8, OP_RESAMPLE, OP_WINATTN = 5, 6, 7, 10, 11, 12, 13
SEG_F16, SEG_E5M2 = 0, 1
DT_BF16, DT_FP16 = 0, 1
ACT_NONE, ACT_RELU, ACT_GELU = 0, 1, 2
NOISE_POISSON, NOISE_GAUSSIAN, NOISE_SALTPEPPER = 1, 2, 3
RNG_INJECTED, RNG_PHILOX = 0, 1

# every symbol include/pssr_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "pssr_last_error": (c_char_p, []),
    "pssr_version": (c_char_p, []),
    "pssr_launch_count": (c_int64, []),
    "pssr_debug_trace": (c_int32, [c_void_p, c_int64]),
    "pssr_normalize_resized_workspace_bytes": (c_int64, [c_int32]),
    "pssr_normalize_preds_resized": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_double,
                                               c_double, c_void_p, c_void_p]),
    "pssr_profile_hist": (c_int32, [c_void_p, c_int32, c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "pssr_tiff_probe": (c_int32, [c_char_p, POINTER(c_int32), POINTER(c_int32), POINTER(c_int32), POINTER(c_int32), POINTER(c_int32)]),
    "pssr_tiff_read": (c_int32, [c_char_p, c_void_p, c_int64]),
    "pssr_tiff_write": (c_int32, [c_char_p, c_void_p, c_int32, c_int32, c_int32, c_int32]),
    "pssr_crappify": (c_int32, [POINTER(CrappifyArgs), c_void_p]),
    "pssr_table_fetch": (c_int32, [c_void_p, c_void_p, c_int64, c_void_p]),
    "pssr_noise_chain": (c_int32, [c_void_p, c_int32, c_void_p, c_int64, POINTER(NoiseStage), c_int32, c_int32, c_uint64, c_void_p]),
    "pssr_resize_bilinear": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "pssr_plan_create": (c_int32, [POINTER(Op), c_int32, c_int32, POINTER(c_void_p)]),
    "pssr_plan_run": (c_int32, [c_void_p, c_void_p]),
    "pssr_plan_run_range": (c_int32, [c_void_p, c_int32, c_int32, c_void_p]),
    "pssr_plan_num_ops": (c_int32, [c_void_p]),
    "pssr_plan_destroy": (None, [c_void_p]),
    "pssr_stitch": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "pssr_metric_workspace_bytes": (c_int64, [c_int32, c_int32, c_int32]),
    "pssr_metric_sums": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p]),
    "pssr_normalize_workspace_bytes": (c_int64, [c_int32]),
    "pssr_normalize_preds": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32,
                                       c_double, c_double, c_void_p, c_void_p]),
}

_lib = None


def lib():
    """The loaded library; raises if it has not been built (no CPU fallback exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(pssr2_b200 has no CPU fallback)")
        l = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().pssr_last_error().decode(errors="replace")
        raise RuntimeError(f"libpssr_b200 {what} failed ({rc}): {msg}")


def current_stream_ptr(device=None):
    """cudaStream_t of torch's current stream on ``device`` (default: the current device).  The C side keys its per-device state
    on cudaGetDevice(), so callers also enter ``on_device(device)`` around every library call."""
    import torch
    return c_void_p(torch.cuda.current_stream(device).cuda_stream)


def on_device(device):
    """Context manager making ``device`` the current CUDA device for the duration of a library call (tensors on cuda:1 must be
    processed with device 1 current: kernels launch into the current device's context)."""
    import torch
    return torch.cuda.device(device)


def same_device(*tensors):
    """The common device of the given CUDA tensors (None entries ignored); inputs that span devices are rejected."""
    devs = {t.device for t in tensors if t is not None}
    if len(devs) != 1:
        raise ValueError(f"all tensors of one call must live on one CUDA device, got {sorted(str(d) for d in devs)}")
    return next(iter(devs))


def launch_count() -> int:
    return int(lib().pssr_launch_count())
