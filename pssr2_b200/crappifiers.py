"""Crappifier classes with the reference's interface (pssr/crappifiers.py:6-105), backed by the
CUDA noise chain in ``csrc/crappify.cu``.

``crappify(image: np.ndarray) -> np.ndarray`` keeps the reference contract (noise only, no
downscale; works on any float array).  Datasets do not call it per tile: they ask each crappifier
for its resolved ``NoiseSpec`` and run tile extraction + downscale + noise as ONE fused launch
(``ops.crappify``).  On-device randomness is counter-based Philox keyed by (seed, tile, pixel); the
per-call seed is drawn from NumPy's global legacy generator, so ``np.random.seed(n)`` makes a run
reproducible exactly as it does for the reference (the streams themselves differ from MT19937:
bit-exactness with the reference is defined with injected draws, see tests/).
"""
from abc import ABC, abstractmethod

import numpy as np
import torch

from . import _lib, ops
from ._lib import NOISE_GAUSSIAN, NOISE_POISSON, NOISE_SALTPEPPER


def _fresh_seed() -> int:
    return int(np.random.randint(0, 2 ** 31 - 1)) | (int(np.random.randint(0, 2 ** 31 - 1)) << 31)


class Crappifier(ABC):
    r"""Crappifier base class for custom crappifiers. Override :meth:`crappify` for logic."""

    @abstractmethod
    def crappify(self, image: np.ndarray):
        raise NotImplementedError('"crappify" method not implemented.')

    def __call__(self, image: np.ndarray):
        return self.crappify(image)

    # --- fused-path hooks -------------------------------------------------------------
    def noise_specs(self):
        """Resolved device noise stages for ONE crappify call (draws ``spread`` like the reference), or
        None for a custom crappifier that only provides a host ``crappify``."""
        return None

    clip_between = False

    def has_spread(self):
        """True if the crappifier redraws its intensity on every call (spread > 0): datasets then resolve it per item."""
        return float(getattr(self, "spread", 0) or 0) > 0


def _resolve_intensity(intensity, spread):
    # max(np.random.normal(intensity, spread), 0) if spread > 0 else intensity  (crappifiers.py:63,85,104)
    if spread > 0:
        return max(np.random.normal(intensity, spread), 0), False
    return intensity, True


class _DeviceCrappifier(Crappifier):
    def crappify(self, image: np.ndarray, _specs=None):
        image = np.asarray(image)
        specs = self.noise_specs() if _specs is None else _specs
        x = torch.as_tensor(np.ascontiguousarray(image if image.dtype in (np.float32, np.float64) else image.astype(np.float64)))
        out = run_noise_chain(x.cuda(), specs, self.clip_between, _fresh_seed())
        res = out.cpu().numpy().reshape(image.shape)
        return res.astype(np.float32) if self._returns_f32(specs, image.dtype) else res

    @staticmethod
    def _returns_f32(specs, in_dtype):
        # NumPy dtype flow of the reference: Poisson / AdditiveGaussian promote to float64, SaltPepper yields float32
        return specs[-1].kind == NOISE_SALTPEPPER


def run_noise_chain(x: torch.Tensor, specs, clip_between, seed):
    import ctypes
    x = x.contiguous()
    out = torch.empty(x.shape, dtype=torch.float64, device=x.device)
    arr = (_lib.NoiseStage * len(specs))()
    for i, s in enumerate(specs):
        arr[i].kind, arr[i].rng = s.kind, _lib.RNG_PHILOX
        arr[i].intensity, arr[i].gain, arr[i].mix_in_f32 = s.intensity, s.gain, 1 if s.mix_in_f32 else 0
        arr[i].injected = None
    with _lib.on_device(x.device):
        _lib.check(_lib.lib().pssr_noise_chain(x.data_ptr(), 1 if x.dtype == torch.float64 else 0, out.data_ptr(), x.numel(), arr,
                                               len(specs), 1 if clip_between else 0, seed, _lib.current_stream_ptr(x.device)),
                   "pssr_noise_chain")
    return out


class MultiCrappifier(_DeviceCrappifier):
    def __init__(self, *args, clip: bool = True):
        r"""Chains multiple crappifiers sequentially (pssr/crappifiers.py:26-43)."""
        self.crappifiers = args
        self.clip = clip

    @property
    def clip_between(self):
        return self.clip

    def has_spread(self):
        return any(isinstance(c, Crappifier) and c.has_spread() for c in self.crappifiers)

    def _device_chain(self):
        """True if the chain can run as one device noise chain: device crappifiers only; a nested MultiCrappifier must clip
        like this one (the kernel has ONE clip-between flag per chain)."""
        for c in self.crappifiers:
            if not isinstance(c, _DeviceCrappifier):
                return False
            if isinstance(c, MultiCrappifier) and (c.clip != self.clip or not c._device_chain()):
                return False
        return True

    def noise_specs(self):
        """One resolution of every stage, in chain order (each stage draws its ``spread`` exactly once per call)."""
        if not self._device_chain():
            return None
        specs = []
        for c in self.crappifiers:
            specs += c.noise_specs()
        return specs

    def crappify(self, image: np.ndarray):
        specs = self.noise_specs()          # resolved ONCE: a second call would consume extra np.random draws with spread > 0
        if specs is None or len(specs) > 4:  # custom host crappifier / nested clip / long chain: chain on the host like the reference
            for c in self.crappifiers:
                image = c.crappify(image)
                if self.clip:
                    image = np.clip(image, 0, 255)
            return image
        return super().crappify(image, _specs=specs)


class AdditiveGaussian(_DeviceCrappifier):
    def __init__(self, intensity: float = 13, gain: float = 0, spread: float = 0):
        r"""Additive Gaussian noise: std ``intensity``, mean ``gain`` (pssr/crappifiers.py:45-64)."""
        self.intensity, self.gain, self.spread = intensity, gain, spread

    def noise_specs(self):
        i, _ = _resolve_intensity(self.intensity, self.spread)
        return [ops.NoiseSpec(NOISE_GAUSSIAN, i, self.gain)]


class Poisson(_DeviceCrappifier):
    def __init__(self, intensity: float = 1, gain: float = 0, spread: float = 0):
        r"""Poisson (shot) noise mixed in with weight ``intensity`` (pssr/crappifiers.py:66-86)."""
        self.intensity, self.gain, self.spread = intensity, gain, spread

    def noise_specs(self):
        i, python_scalar = _resolve_intensity(self.intensity, self.spread)
        # x*(1-i): float32 when i is a python scalar, float64 when it is the np.float64 spread draw (NumPy>=2)
        return [ops.NoiseSpec(NOISE_POISSON, i, self.gain, mix_in_f32=python_scalar and not isinstance(i, np.floating))]


class SaltPepper(_DeviceCrappifier):
    def __init__(self, intensity: float = 0.5, gain: float = 0, spread: float = 0):
        r"""Salt and pepper noise on ``intensity`` percent of the values (pssr/crappifiers.py:88-105)."""
        self.intensity, self.gain, self.spread = intensity / 100, gain, spread

    def noise_specs(self):
        i, _ = _resolve_intensity(self.intensity, self.spread)
        return [ops.NoiseSpec(NOISE_SALTPEPPER, i, self.gain)]
