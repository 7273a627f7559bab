"""Host mirror of the stitch / normalise / metric helpers (pssr/util.py:54-231) on the CUDA kernels."""
import glob
import inspect
import math
import os
from pathlib import Path

import numpy as np
import torch

from . import ops


def _force_list(item):
    """pssr/util.py:220-226."""
    if type(item) is not list:
        try:
            return list(item)
        except Exception:
            return [item]
    return item


def _get_callbacks(raw):
    """pssr/util.py:228-231: a callback taking exactly one non-self argument receives locals()."""
    callbacks = [] if raw is None else _force_list(raw)
    callback_locals = [len([arg for arg in inspect.getfullargspec(cb).args if arg != "self"]) == 1 for cb in callbacks]
    return callbacks, callback_locals


def pixel_metric(mse: float, image_range: int = 255):
    r"""Average pixel error from a mean squared error (pssr/util.py:207-215)."""
    return math.sqrt(mse) * image_range


def _sort_tiles(name: str):
    """pssr/util.py:110-114: sort key (slice, tile) from ``name_{tile}_{slice}[.ext]``."""
    if "." not in name:
        name += "."
    parts = name.replace(".", "_").split("_")
    return int(parts[-2]), int(parts[-3])


def normalize_preds(hr, hr_hat, pmin: float = 0.1, pmax: float = 99.9):
    r"""Normalizes prediction intensities to ground truth (pssr/util.py:139-191) on the device.

    Accepts uint8 arrays / tensors of identical shape [..., H, W]; returns uint8 arrays of the input shapes
    (NumPy in -> NumPy out, CUDA tensors in -> CUDA tensors out)."""
    as_numpy = not isinstance(hr, torch.Tensor)
    a = torch.as_tensor(np.asarray(hr)) if as_numpy else hr
    b = torch.as_tensor(np.asarray(hr_hat)) if not isinstance(hr_hat, torch.Tensor) else hr_hat
    if a.dim() != b.dim():
        raise ValueError(f"hr and hr_hat must have the same number of dimensions. Dimension lengths are {tuple(a.shape)} and {tuple(b.shape)} respectively.")
    sa, sb = a.shape, b.shape
    if a.dim() < 3:
        a, b = a[None], b[None]
    a, b = a.reshape(-1, *a.shape[-2:]), b.reshape(-1, *b.shape[-2:])
    if len(a) != len(b):
        raise ValueError(f"hr and hr_hat must have the same number of images. Received {len(a)} and {len(b)} images respectively.")
    if a.dtype != torch.uint8 or b.dtype != torch.uint8:
        raise TypeError("the device normalize_preds expects uint8 images (what `_pred_array` produces, predict.py:245-246)")
    if a.shape != b.shape:
        # util.py:179: hr_hat is enlarged to hr's grid for the covariance only (skimage.transform.resize); enlarging is what
        # `_collage_preds` needs -- shrinking would add skimage's Gaussian anti-aliasing filter and is not on this path
        if b.shape[-2] > a.shape[-2] or b.shape[-1] > a.shape[-1]:
            raise NotImplementedError("normalize_preds with hr_hat larger than hr (anti-aliased skimage.transform.resize) is off the accelerated path")
        oa, ob = ops.normalize_preds_resized_u8(a.cuda(), b.cuda(), pmin, pmax)
    else:
        oa, ob = ops.normalize_preds_u8(a.cuda(), b.cuda(), pmin, pmax)
    oa, ob = oa.reshape(sa), ob.reshape(sb)
    if as_numpy:
        return oa.cpu().numpy(), ob.cpu().numpy()
    return oa, ob


def _sheet_shapes(lr_path):
    """{sheet name: (frames, H, W)} from a directory of .tif sheets (Pillow), a dict name -> shape/array, or a
    dataset from pssr2_b200.data (in-memory sheets)."""
    if isinstance(lr_path, dict):
        return {k: (tuple(v) if isinstance(v, (tuple, list)) else tuple(np.asarray(v).shape)) for k, v in lr_path.items()}
    if hasattr(lr_path, "hr_files") and hasattr(lr_path, "_shapes"):
        return {n.split("/")[-1].split(".")[0]: (f,) + tuple(hw) for n, f, hw in zip(lr_path.hr_files, lr_path._frames_total, lr_path._shapes)}
    files = glob.glob(f"{lr_path}/*.tif", recursive=True)
    if len(files) == 0:
        raise FileExistsError("No files exist in lr_path.")
    from .io import tiff_probe
    return {f.split("/")[-1].split(".")[0]: tiff_probe(f)[:3] for f in files}


def reassemble_sheets(pred_path, lr_path, lr_scale: int, overlap: int = 0, margin: int = 0, out_dir: str = "sheets"):
    r"""Reassembles image sheets from predicted tiles (pssr/util.py:54-108) with the CUDA stitch kernel.

    ``pred_path``: dict of named tiles from :func:`predict_images` (or a directory of tile TIFFs), or the device-resident
    ``TilePreds`` of ``predict_images(..., keep_on_device=True)`` -- then no tile visits the host, only the stitched sheets do;
    ``lr_path``: directory of low-resolution sheets, or -- for in-memory work -- a dict
    ``{sheet_name: shape}`` / the dataset itself.  Returns the list of uint8 sheets if ``out_dir`` is None (sheets without
    tiles in ``pred_path`` -- another rank's share -- are skipped, as in the reference)."""
    if margin > overlap:
        raise ValueError(f"The value of margin cannot be greater than overlap. Given {margin} and {overlap} respectively.")
    shapes = _sheet_shapes(lr_path)
    if out_dir:
        os.makedirs(out_dir, exist_ok=True)
    outs = []
    for sheet, lr_shape in shapes.items():
        if hasattr(pred_path, "device_tiles"):
            files = sorted([f for f in pred_path.keys() if "_".join(f.split("_")[:-2]) == sheet], key=_sort_tiles)
            if len(files) == 0:
                continue
            batched = pred_path.device_tiles(files)
        elif isinstance(pred_path, dict):
            files = sorted([f for f in pred_path.keys() if "_".join(f.split("_")[:-2]) == sheet], key=_sort_tiles)
            tiles = [pred_path[f] for f in files]
            if len(tiles) == 0:
                continue
            if isinstance(tiles[0], torch.Tensor):
                batched = torch.stack([t.reshape(t.shape[-2:]) for t in tiles]).cuda()
            else:
                batched = torch.as_tensor(np.asarray([np.asarray(t).squeeze() for t in tiles])).cuda()
        else:
            from .io import read_tiff
            files = sorted(glob.glob(f"{pred_path}/{sheet}*"), key=_sort_tiles)
            batched = torch.as_tensor(np.asarray([read_tiff(f).squeeze() for f in files])).cuda()
        T = batched.shape[1]
        n_rows = (lr_shape[1] * lr_scale - T) // (T - overlap * lr_scale) + 1          # util.py:96
        n_cols = (lr_shape[2] * lr_scale - batched.shape[2]) // (batched.shape[2] - overlap * lr_scale) + 1
        stacks = batched.shape[0] // n_rows // n_cols
        image = ops.stitch(batched[:stacks * n_rows * n_cols], n_rows, n_cols, overlap * lr_scale, margin).cpu().numpy()
        if out_dir:
            from .io import write_tiff
            write_tiff(f"{out_dir}/{sheet}.tif", image)
        else:
            outs.append(image)
    if out_dir is None:
        return outs
