"""CPU restatement of the non-network stages of the PSSR2 test/predict path.
TEST INFRASTRUCTURE ONLY.  NumPy semantics are those of NumPy >= 2 (NEP 50 promotion), the version
installed wherever this oracle can run; the reference pins ``numpy ^1.22.4`` (pyproject.toml:26).

Pinned against the reference's own functions (imported through oracle/refshim.py) in
tests/test_oracle_vs_reference.py and through the committed vectors in tests/golden/.
"""
import math

import numpy as np

from .pillow_resize import resize_bilinear
from . import thirdparty as tp


# --------------------------------------------------------------------------- tiling (a-1, a-2)
def n_tiles(shape_hw, size, stride):
    """pssr/data.py:682-687."""
    x, y = shape_hw
    return max(0, (x - size) // stride + 1), max(0, (y - size) // stride + 1)


def sliding_window(sheet, size, stride, n_frames, n_slices, idx, slide=False):
    """pssr/data.py:629-660: flat tile/slice index -> [f, size, size] view."""
    _, tiles_y = n_tiles(sheet.shape[-2:], size, stride)
    tile_idx = idx // n_slices
    sx = tile_idx // tiles_y * stride
    sy = tile_idx % tiles_y * stride
    img = sheet[..., sx:sx + size, sy:sy + size]
    if n_frames is None:
        return img
    f = idx % n_slices
    if not slide:
        f *= n_frames
    return img[f:f + n_frames]


def square_crop(image, max_res):
    """pssr/data.py:536-546."""
    h, w = image.shape[-2:]
    if [h, w] == [max_res] * 2:
        return image
    size = min(h, w, max_res)
    sx, sy = (h - size) // 2, (w - size) // 2
    return image[:, sx:sx + size, sy:sy + size]


def pad_image(image, res):
    """pssr/data.py:548-551 (pad amount from the LAST dim, bottom/right, reflect)."""
    if image.shape[-1] < res:
        p = res - image.shape[-1]
        return np.stack([np.pad(c, [[0, p]] * 2, mode="reflect") for c in image])
    return image


def slice_center(image, n_frames):
    """pssr/data.py:662-668."""
    center = image.shape[-3] // 2
    half = n_frames // 2
    if n_frames % 2 == 0:
        return image[..., center - half:center + half, :, :]
    return image[..., center - half:center + half + 1, :, :]


# ------------------------------------------------------------------------ crappifiers (a-5..a-8)
class Injected:
    """Recorded noise draws for one crappify call, consumed in stage order."""

    def __init__(self, draws):
        self.draws = list(draws)
        self.pos = 0

    def next(self):
        d = self.draws[self.pos]
        self.pos += 1
        return d


def poisson_stage(image, y, intensity=1, gain=0):
    """pssr/crappifiers.py:82-86 with the draw ``y = np.random.poisson(clip(image,0,inf))`` injected."""
    x = image.astype(np.float32)
    return x * (1 - intensity) + y * intensity + gain


def gaussian_stage(image, g):
    """pssr/crappifiers.py:62-64 with ``g = np.random.normal(gain, intensity, shape)`` injected."""
    return image.astype(np.float32) + g


def saltpepper_stage(image, flipped, salted, gain=0):
    """pssr/crappifiers.py:103-105 (+ skimage random_noise s&p) with both Bernoulli masks injected."""
    v = np.clip(image.astype(np.float32) + gain, 0, 255) / 255
    return tp.random_noise(v, mode="s&p", flipped=flipped, salted=salted) * 255


def crappify_chain(lr, stages, clip=True):
    """MultiCrappifier.crappify (pssr/crappifiers.py:38-43).  stages: list of tuples
    ("poisson", y, intensity, gain) | ("gaussian", g) | ("saltpepper", flipped, salted, gain)."""
    for st in stages:
        if st[0] == "poisson":
            lr = poisson_stage(lr, *st[1:])
        elif st[0] == "gaussian":
            lr = gaussian_stage(lr, *st[1:])
        elif st[0] == "saltpepper":
            lr = saltpepper_stage(lr, *st[1:])
        else:
            raise ValueError(st[0])
        if clip:
            lr = np.clip(lr, 0, 255)
    return lr


def gen_pair(hr, hr_res, lr_scale, stages, n_frames=None, clip=True, multi=True, rotation=False):
    """pssr/data.py:471-495 without transforms.  ``stages is None`` <=> crappifier=None.
    ``multi=False`` models a bare (non-Multi) crappifier: no clip after the stage.
    ``rotation``: False or [rot90?, flip axis 1 | 2 | (1, 2)] as drawn at pssr/data.py:108 (applied at :478-480)."""
    hr = pad_image(square_crop(hr, hr_res), hr_res)
    if rotation:
        hr = np.rot90(hr, axes=(1, 2)) if rotation[0] else hr
        hr = np.flip(hr, axis=rotation[1])
    lr = resize_bilinear(np.ascontiguousarray(hr), hr_res // lr_scale, hr_res // lr_scale).astype(np.float32)
    if stages is not None:
        lr = crappify_chain(lr, stages, clip=clip and multi)
        lr = np.clip(lr.round(), 0, 255)
    if n_frames is not None and n_frames[0] != n_frames[1]:
        if not n_frames[1] > hr.shape[-3]:
            hr = slice_center(hr, n_frames[1])
        if not n_frames[0] > lr.shape[-3]:
            lr = slice_center(lr, n_frames[0])
    return hr.astype(np.float32), lr.astype(np.float32)


# ------------------------------------------------------------------------------ glue (a-13)
def pred_array(data, n_frames=1):
    """pssr/predict.py:245-246: clip -> uint8 truncation -> centre channel."""
    return slice_center(np.clip(np.asarray(data), 0, 255).astype(np.uint8), n_frames)


# --------------------------------------------------------------------------- normalize (a-14)
def normalize_minmax(x, pmin=0.1, pmax=99.9, eps=1e-20, dtype=np.float32):
    """pssr/util.py:193-205."""
    x_min = np.percentile(x, pmin, keepdims=True)
    x_max = np.percentile(x, pmax, keepdims=True)
    x = x.astype(dtype, copy=False)
    x_min = x_min.astype(dtype, copy=False)
    x_max = x_max.astype(dtype, copy=False)
    return (x - x_min) / (x_max - x_min + dtype(eps))


def normalize_preds(hr, hr_hat, pmin=0.1, pmax=99.9):
    """pssr/util.py:139-191; differing resolutions go through ``oracle.thirdparty.resize`` (util.py:179)."""
    from .thirdparty import resize
    hr, hr_hat = np.asarray(hr), np.asarray(hr_hat)
    hr_shape, hh_shape = hr.shape, hr_hat.shape
    hr = hr.reshape(-1, *hr.shape[-2:])
    hr_hat = hr_hat.reshape(-1, *hr_hat.shape[-2:])
    outs_a, outs_b = [], []
    for a, b in zip(hr, hr_hat):
        a = a.astype(np.float32)
        b = b.astype(np.float32)
        base_max = np.percentile(a, pmax)
        base_mean = np.mean(a)
        a = normalize_minmax(a, pmin, pmax)
        b = b - np.mean(b)
        a = a - np.mean(a)
        scaled = resize(b, a.shape) if b.shape != a.shape else b
        amp = np.cov(scaled.flatten(), a.flatten())[0, 1] / np.var(b.flatten())
        b = amp * b
        a, b = (a - a.min()) * base_max, (b - a.min()) * base_max
        a, b = a / (a.mean() / base_mean), b / (b.mean() / base_mean)
        outs_a.append(a)
        outs_b.append(b)
    A, Bm = np.asarray(outs_a).clip(0, 255), np.asarray(outs_b).clip(0, 255)
    return A.reshape(hr_shape).astype(np.uint8), Bm.reshape(hh_shape).astype(np.uint8)


# ----------------------------------------------------------------------------- metrics (a-15)
def image_metrics(hr, hr_hat, image_range=255):
    """pssr/predict.py:193-203 for one uint8 pair [1,H,W]: (mse, pixel, psnr, ssim)."""
    mse = float(np.mean((hr / image_range - hr_hat / image_range) ** 2))
    pixel = math.sqrt(mse) * image_range                       # pssr/util.py:207-215
    psnr = tp.peak_signal_noise_ratio(hr, hr_hat, data_range=image_range)
    ssim = tp.structural_similarity(hr.squeeze(), hr_hat.squeeze(), data_range=image_range)
    return mse, pixel, psnr, ssim


# ------------------------------------------------------------------------------ stitch (a-16)
def patch_images(batched, n_cols, n_rows, overlap, margin):
    """pssr/util.py:116-137 (float64 sum / count; caller casts to uint8 by truncation, :100)."""
    T = batched.shape[-1]
    step = T - overlap
    H, W = n_rows * step + overlap, n_cols * step + overlap
    collage = np.zeros((H, W))
    count = np.zeros((H, W))
    for idx in range(n_rows * n_cols):
        row, col = idx // n_cols, idx % n_cols
        sr, sc = row * step, col * step
        m = [margin if row != 0 else 0, margin if row != n_rows - 1 else 0,
             margin if col != 0 else 0, margin if col != n_cols - 1 else 0]
        collage[sr + m[0]:sr + T - m[1], sc + m[2]:sc + T - m[3]] += batched[idx, m[0]:T - m[1], m[2]:T - m[3]]
        count[sr + m[0]:sr + T - m[1], sc + m[2]:sc + T - m[3]] += 1
    count[count == 0] = 1
    return collage / count


def stitch_sheets(tiles, n_rows, n_cols, overlap, margin):
    """tiles [stacks*n_rows*n_cols, T, T] uint8 -> uint8 [stacks, H, W] (pssr/util.py:96-100)."""
    per = n_rows * n_cols
    stacks = tiles.shape[0] // per
    return np.asarray([patch_images(tiles[i * per:(i + 1) * per], n_cols, n_rows, overlap, margin)
                       for i in range(stacks)], dtype=np.uint8)
