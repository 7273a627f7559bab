"""NumPy restatement of Pillow's ``Image.resize(size, BILINEAR)`` for modes ``L`` (uint8) and
``I;16`` (uint16).  TEST INFRASTRUCTURE ONLY.

The reference calls Pillow at pssr/data.py:483 (``Image.fromarray(channel).resize([hr_res//lr_scale]*2,
Image.Resampling.BILINEAR)``) and pssr/train.py:365.  Pillow (pin ``pillow >=9.1.0``, pyproject.toml:27;
12.2.0 installed here) implements it in ``src/libImaging/Resample.c``:

* ``precompute_coeffs``: per output index ``xx`` a window ``[xmin, xmin+xmax)`` of input samples with
  triangle weights ``w = 1 - |(x + xmin - center + 0.5) / filterscale|`` (support = filterscale when
  downscaling), normalised by their sum;
* two passes, HORIZONTAL THEN VERTICAL, the intermediate image rounded to the image dtype;
* 8 bpc: coefficients quantised ``(int)(0.5 + w * 2**22)``, accumulator starts at ``2**21``, result
  ``clip8(acc >> 22)``;
* 16 bpc (``I;16``): double coefficients, sequential double sum (no FMA), ``(int)(ss + 0.5)``, bytes
  clipped separately (never triggers for a normalised non-negative filter).

Pinned bit-for-bit against Pillow itself in tests/test_oracle.py (Pillow is installed in this image).
"""
import numpy as np

PRECISION_BITS = 32 - 8 - 2


def precompute_coeffs(in_size: int, out_size: int):
    """Returns (bounds[out,2] int32 = (xmin, count), kk[out, ksize] float64) like Resample.c."""
    scale = filterscale = in_size / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 1.0 * filterscale  # BILINEAR support = 1.0
    ksize = int(np.ceil(support)) * 2 + 1
    kk = np.zeros((out_size, ksize), np.float64)
    bounds = np.zeros((out_size, 2), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        ww = 0.0
        for x in range(xmax):
            a = (x + xmin - center + 0.5) * ss
            if a < 0:
                a = -a
            w = 1.0 - a if a < 1.0 else 0.0
            kk[xx, x] = w
            ww += w
        if ww != 0.0:
            for x in range(xmax):
                kk[xx, x] /= ww
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def quantize_coeffs_8bpc(kk):
    """normalize_coeffs_8bpc: (int)(+-0.5 + k * 2**22), C truncation toward zero."""
    v = kk * float(1 << PRECISION_BITS)
    return np.where(kk < 0, np.trunc(-0.5 + v), np.trunc(0.5 + v)).astype(np.int64)


def _pass_u8(img, bounds, kq, axis):
    """One 8 bpc pass along `axis` (1 = horizontal, 0 = vertical)."""
    img = np.moveaxis(img, axis, -1).astype(np.int64)
    out = np.empty(img.shape[:-1] + (len(bounds),), np.uint8)
    for xx, (xmin, cnt) in enumerate(bounds):
        acc = (img[..., xmin:xmin + cnt] * kq[xx, :cnt]).sum(-1) + (1 << (PRECISION_BITS - 1))
        out[..., xx] = np.clip(acc >> PRECISION_BITS, 0, 255)
    return np.moveaxis(out, -1, axis)


def _pass_u16(img, bounds, kk, axis):
    """One 16 bpc pass: sequential double accumulation in tap order, ROUND_UP = (int)(ss + 0.5)."""
    img = np.moveaxis(img, axis, -1).astype(np.float64)
    out = np.empty(img.shape[:-1] + (len(bounds),), np.uint16)
    for xx, (xmin, cnt) in enumerate(bounds):
        ss = np.zeros(img.shape[:-1], np.float64)
        for x in range(cnt):
            ss = ss + img[..., xmin + x] * kk[xx, x]
        ss_int = np.trunc(ss + 0.5).astype(np.int64)
        lo = ss_int % 256
        hi = np.clip(ss_int >> 8, 0, 255)
        out[..., xx] = (lo + (hi << 8)).astype(np.uint16)
    return np.moveaxis(out, -1, axis)


def resize_bilinear(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """img [..., H, W] uint8 or uint16 -> [..., out_h, out_w], same dtype, Pillow-exact."""
    h, w = img.shape[-2:]
    out = img
    if img.dtype == np.uint8:
        if out_w != w:
            b, k = precompute_coeffs(w, out_w)
            out = _pass_u8(out, b, quantize_coeffs_8bpc(k), -1)
        if out_h != h:
            b, k = precompute_coeffs(h, out_h)
            out = _pass_u8(out, b, quantize_coeffs_8bpc(k), -2)
    elif img.dtype == np.uint16:
        if out_w != w:
            b, k = precompute_coeffs(w, out_w)
            out = _pass_u16(out, b, k, -1)
        if out_h != h:
            b, k = precompute_coeffs(h, out_h)
            out = _pass_u16(out, b, k, -2)
    else:
        raise TypeError(f"unsupported dtype {img.dtype} (Pillow modes L and I;16 only)")
    return out


def pillow_resize(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """The real thing (used to pin the restatement): per-channel Pillow BILINEAR resize."""
    from PIL import Image
    flat = img.reshape((-1,) + img.shape[-2:])
    res = np.stack([np.asarray(Image.fromarray(c).resize((out_w, out_h), Image.Resampling.BILINEAR)) for c in flat])
    return res.reshape(img.shape[:-2] + (out_h, out_w))
