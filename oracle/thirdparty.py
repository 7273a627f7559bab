"""Restatements of the third-party routines the reference calls on the hot path and that
are NOT installed in this environment (scikit-image, timm).  TEST INFRASTRUCTURE ONLY.

parity unpinned: these follow the published scikit-image 0.2x / timm 1.x algorithms; the
real packages are absent, so they are checked only against brute-force formulas in
``tests/test_oracle.py``.

Reference call sites:
  skimage.metrics.peak_signal_noise_ratio / structural_similarity  pssr/predict.py:5,201,203
  skimage.util.random_noise(mode="s&p")                            pssr/crappifiers.py:3,105
  timm.layers.LayerNorm2d / EffectiveSEModule / DropPath           pssr/models/_rdnet.py:11
  timm.models.named_apply                                          pssr/models/_rdnet.py:12
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
from scipy.ndimage import uniform_filter


# ----------------------------------------------------------------------------- skimage
def peak_signal_noise_ratio(image_true, image_test, *, data_range=None):
    """skimage.metrics.peak_signal_noise_ratio: 10*log10(R^2 / mean((a-b)^2)) in float64."""
    a = np.asarray(image_true).astype(np.float64)
    b = np.asarray(image_test).astype(np.float64)
    err = np.mean((a - b) ** 2, dtype=np.float64)
    with np.errstate(divide="ignore"):
        return float(10 * np.log10((data_range ** 2) / err))


def structural_similarity(im1, im2, *, data_range=None, win_size=7, K1=0.01, K2=0.03):
    """skimage.metrics.structural_similarity defaults: uniform 7x7 window, sample
    covariance (cov_norm = NP/(NP-1)), float64, mean over the image cropped by 3 px."""
    im1 = np.asarray(im1).astype(np.float64)
    im2 = np.asarray(im2).astype(np.float64)
    if im1.shape != im2.shape:
        raise ValueError("Input images must have the same dimensions.")
    if min(im1.shape) < win_size:
        raise ValueError("win_size exceeds image extent.")
    NP = win_size ** im1.ndim
    cov_norm = NP / (NP - 1)
    ux = uniform_filter(im1, size=win_size)
    uy = uniform_filter(im2, size=win_size)
    uxx = uniform_filter(im1 * im1, size=win_size)
    uyy = uniform_filter(im2 * im2, size=win_size)
    uxy = uniform_filter(im1 * im2, size=win_size)
    vx = cov_norm * (uxx - ux * ux)
    vy = cov_norm * (uyy - uy * uy)
    vxy = cov_norm * (uxy - ux * uy)
    R = data_range
    C1 = (K1 * R) ** 2
    C2 = (K2 * R) ** 2
    A1, A2, B1, B2 = (2 * ux * uy + C1, 2 * vxy + C2, ux ** 2 + uy ** 2 + C1, vx + vy + C2)
    S = (A1 * A2) / (B1 * B2)
    pad = (win_size - 1) // 2
    sl = tuple(slice(pad, s - pad) for s in S.shape)
    return float(S[sl].mean(dtype=np.float64))


def random_noise(image, mode="s&p", amount=0.05, rng=None, flipped=None, salted=None):
    """skimage.util.random_noise, mode="s&p" only (salt_vs_pepper=0.5, clip=True).

    ``flipped``/``salted`` inject the two Bernoulli masks so the GPU kernel can be
    checked bit-exactly; when absent they are drawn the way scikit-image does
    (fresh ``default_rng``: ``rng.random(shape) <= p``)."""
    if mode != "s&p":
        raise NotImplementedError(mode)
    image = np.asarray(image)
    if image.dtype not in (np.float32, np.float64):
        image = image.astype(np.float64)
    low_clip = -1.0 if image.min() < 0 else 0.0
    out = image.copy()
    rng = np.random.default_rng(rng)
    if flipped is None:
        flipped = rng.random(out.shape) <= amount
    if salted is None:
        salted = rng.random(out.shape) <= 0.5
    out[flipped & salted] = 1
    out[flipped & ~salted] = low_clip
    return np.clip(out, low_clip, 1.0)


def resize(image, output_shape, **kw):
    """skimage.transform.resize(image, output_shape) with its defaults (order 1, mode "reflect", clip, anti_aliasing only when an
    axis shrinks), as called at pssr/util.py:179.  scikit-image >= 0.19 implements it as scipy.ndimage.zoom(order=1, mode="mirror",
    grid_mode=True) -- scipy IS installed, so the interpolation itself is the library's; the wrapper around it is parity-unpinned
    (scikit-image absent).  Shrinking (Gaussian anti-aliasing prefilter) is not restated."""
    from scipy import ndimage as ndi
    image = np.asarray(image)
    if kw:
        raise NotImplementedError("only the default arguments are restated")
    if any(o < i for o, i in zip(output_shape, image.shape)):
        raise NotImplementedError("anti-aliased shrinking is not restated")
    if image.dtype == np.float16:
        image = image.astype(np.float32)
    zoom = [o / i for o, i in zip(output_shape, image.shape)]
    out = ndi.zoom(image, zoom, order=1, mode="mirror", cval=0, grid_mode=True)
    return np.clip(out, image.min(), image.max())          # clip=True: keep the input range


def gaussian(image, sigma, channel_axis=None):  # Blur crappifier, out of scope
    raise NotImplementedError("skimage.filters.gaussian is out of scope (Blur)")


# -------------------------------------------------------------------------------- timm
class LayerNorm2d(nn.LayerNorm):
    """timm.layers.LayerNorm2d: LayerNorm over C of an NCHW tensor, eps=1e-6."""

    def __init__(self, num_channels, eps=1e-6, affine=True):
        super().__init__(num_channels, eps=eps, elementwise_affine=affine)

    def forward(self, x):
        x = x.permute(0, 2, 3, 1)
        x = F.layer_norm(x, self.normalized_shape, self.weight, self.bias, self.eps)
        return x.permute(0, 3, 1, 2)


class EffectiveSEModule(nn.Module):
    """timm.layers.EffectiveSEModule: x * hard_sigmoid(fc(mean_HW(x))), fc = 1x1 conv."""

    def __init__(self, channels, add_maxpool=False, gate_layer="hard_sigmoid", **_):
        super().__init__()
        self.add_maxpool = add_maxpool
        self.fc = nn.Conv2d(channels, channels, kernel_size=1, padding=0)

    def forward(self, x):
        x_se = x.mean((2, 3), keepdim=True)
        if self.add_maxpool:
            x_se = 0.5 * x_se + 0.5 * x.amax((2, 3), keepdim=True)
        x_se = self.fc(x_se)
        return x * (F.relu6(x_se + 3.0) / 6.0)


class DropPath(nn.Module):
    """timm.layers.DropPath: identity in eval mode (the only mode on the predict path)."""

    def __init__(self, drop_prob=0.0, scale_by_keep=True):
        super().__init__()
        self.drop_prob = drop_prob
        self.scale_by_keep = scale_by_keep

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
        if keep > 0.0 and self.scale_by_keep:
            mask.div_(keep)
        return x * mask


def named_apply(fn, module, name="", depth_first=True, include_root=False):
    """timm.models.named_apply."""
    if not depth_first and include_root:
        fn(module=module, name=name)
    for child_name, child in module.named_children():
        child_name = ".".join((name, child_name)) if name else child_name
        named_apply(fn=fn, module=child, name=child_name, depth_first=depth_first, include_root=True)
    if depth_first and include_root:
        fn(module=module, name=name)
    return module
