"""Device-resident mirrors of the reference datasets' predict path (pssr/data.py).

``SlidingDataset`` (pssr/data.py:132-266) and ``ImageDataset`` (:12-130) keep the reference's
constructor arguments, index math (``_get_image_idx`` :697-706, ``_get_val_idx`` :708-730,
``_n_tiles`` :682-687, ``_sliding_window`` :629-638, ``_slice_image`` :649-660) and the attributes
``predict_images`` / ``test_metrics`` read (``val_idx``, ``is_lr``, ``crop_res``, ``lr_scale``,
``hr_res``, ``n_frames``, ``_get_name``).  What changes is where the work happens: sheets are
uploaded to HBM once, and ``batch(indices)`` produces a whole batch of (HR, LR) tiles with ONE fused
CUDA launch (tile gather, crop/reflect-pad, Pillow-exact downscale, noise, round/clip) instead of
per-item NumPy / Pillow work on the host.

Sources: besides a directory path (TIFF / PNG read with Pillow), ``path`` may be a dict
``{name: ndarray}`` or a list of ndarrays ``[frames, H, W]`` / ``[H, W]`` (uint8 or uint16) -- the
benchmark inputs are synthetic in-memory arrays.  File readers for .czi, ``extra_path``, ``transforms``
and training-time rotation are outside the accelerated hot path.
"""
import glob
import os
import random
import warnings
from pathlib import Path

import numpy as np
import torch
from torch.utils.data import Dataset

from . import ops
from .crappifiers import Crappifier, Poisson, _fresh_seed


def _force_list(item):
    if type(item) is not list:
        try:
            return list(item)
        except Exception:
            return [item]
    return item


def _get_n_frames(n_frames):
    """pssr/data.py:689-695 -> None or [in (LR), out (HR)]."""
    if n_frames in [None, -1, [-1]]:
        return None
    n_frames = _force_list(n_frames)
    return n_frames * 2 if len(n_frames) == 1 else n_frames


def _get_image_idx(idx, slices, tiles=None):
    """pssr/data.py:697-706."""
    tiles = [1] * len(slices) if tiles is None else tiles
    image_idx = 0
    for s, t in zip(slices, tiles):
        if idx < s * t:
            return image_idx, idx
        idx -= s * t
        image_idx += 1
    raise IndexError("index out of range")


def _get_val_idx(slices, split, seed, tiles=None):
    """pssr/data.py:708-730 (same use of NumPy's global generator, so the same validation items)."""
    if tiles is not None:
        ts = []
        for s, t in zip(slices, tiles):
            ts.extend([s] * t)
        slices = ts
    val_slices = list(range(len(slices)))
    if seed is not None and split < 1:
        np.random.seed(seed)
        np.random.shuffle(val_slices)
    val_slices = set(val_slices[-max(1, int(split * len(slices))):])
    val_idx, idx = [], 0
    for i, s in enumerate(slices):
        if i in val_slices:
            val_idx.extend(range(idx, idx + s))
        idx += s
    return val_idx


def _n_tiles(shape_hw, size, stride):
    """pssr/data.py:682-687."""
    x, y = shape_hw
    return max(0, (x - size) // stride + 1), max(0, (y - size) // stride + 1)


def _slice_center_range(total, n):
    """pssr/data.py:662-668 as a (start, count) pair."""
    center, half = total // 2, n // 2
    return (center - half, 2 * half) if n % 2 == 0 else (center - half, 2 * half + 1)


def _load_sources(path, extension, stack="TZ", mode="L"):
    """-> (names, arrays [F,H,W] uint8/uint16 or lazy ``io.SheetFile`` sources)."""
    if isinstance(path, dict):
        names, arrays = list(path.keys()), list(path.values())
    elif isinstance(path, (list, tuple)):
        arrays = list(path)
        names = [f"image_{i}" for i in range(len(arrays))]
    else:
        p = Path(path) if type(path) is str else path
        if not path or not p.exists():
            raise FileNotFoundError(f'Path "{p}" does not exist.')
        files = sorted(glob.glob(f"{p}/**/*.{extension}", recursive=True))
        if not len(files) > 0:
            raise FileNotFoundError(f'No .{extension} files exist in path "{p}".')
        # lazy sources (pssr2_b200/io.py): TIFF geometry from a header probe, pixels decoded into pinned memory on a reader
        # thread when a batch first needs the sheet (tifffile.imread keeps the native 8 / 16-bit depth, data.py:621-625);
        # CZI: axis selection / channel mean / max-normalisation to uint8 (data.py:585-619); other formats: Pillow,
        # converted to mode "L" (data.py:640-647)
        from .io import SheetFile
        names = [os.path.relpath(f, p) for f in files]
        arrays = [SheetFile(f, stack=stack, mode=mode) for f in files]
    out = []
    for a in arrays:
        if hasattr(a, "read_pinned"):
            if a.dtype not in (np.uint8, np.uint16) or len(a.shape) != 3:
                raise TypeError(f"{a.path}: images must decode to [frames, H, W] uint8 / uint16")
            out.append(a)
            continue
        if isinstance(a, torch.Tensor):   # e.g. a pinned host stack: uploaded as is (int16 = uint16 container)
            if a.dim() == 2:
                a = a[None]
            if a.dim() != 3 or a.dtype not in (torch.uint8, torch.int16, torch.uint16):
                raise TypeError(f"tensor images must be [frames, H, W] uint8 / uint16, got {tuple(a.shape)} {a.dtype}")
            out.append(a.contiguous())
            continue
        a = np.asarray(a)
        if a.ndim == 2:
            a = a[np.newaxis]
        if a.ndim != 3:
            raise ValueError(f"images must be [frames, H, W] or [H, W], got shape {a.shape}")
        if a.dtype == np.bool_:
            a = a.astype(np.uint8)
        if a.dtype not in (np.uint8, np.uint16):
            raise TypeError(f"images must be uint8 or uint16 (Pillow modes L / I;16), got {a.dtype}")
        out.append(np.ascontiguousarray(a))
    return names, out


class _DeviceDataset(Dataset):
    """Shared machinery: sheets resident in HBM, tile table, fused batch generation.

    Residency.  ``_sources[i]`` is the host side of sheet i (a NumPy array, a -- preferably pinned -- torch tensor, or a lazy
    ``io.SheetFile`` that decodes into a pinned staging buffer on a reader thread); ``_sheets[i]`` is its device copy or None.
    Sheets go up on a side stream the first time a batch touches them, the next one is prefetched while the current one is
    being predicted, and -- when ``max_resident_bytes`` is set -- the least recently used ones are dropped again, so a dataset
    larger than HBM streams through (the reference's ``preload=False`` analogue, pssr/data.py:553-564).  Without a process
    group and with everything fitting the budget the whole dataset is uploaded at construction (``preload=True``)."""

    max_resident_bytes = None      # None: keep every sheet that was uploaded
    preload_bytes = None           # cap on what `preload` uploads at construction (None: everything; the first sheet always)
    auto_resident_frac = 0.5       # a dataset larger than this share of the free HBM gets `max_resident_bytes` set to it

    def _upload(self, arrays, device, preload=True):
        dtypes = {str(a.dtype).replace("torch.", "").replace("int16", "uint16").replace("uuint16", "uint16") for a in arrays}
        if len(dtypes) != 1:
            raise NotImplementedError("all images of one dataset must share their dtype on the device path")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("pssr2_b200 datasets live on CUDA devices (no CPU path); pass device='cuda'")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self._sources = list(arrays)
        self._sheets = [None] * len(arrays)
        self._sheet_events = {}     # sheet index -> upload-complete event not yet waited for by the compute stream
        self._lru = []              # resident sheet indices, least recently used first
        self._upload_stream = torch.cuda.Stream(device=self.device)
        self._frames_total = [a.shape[0] for a in arrays]
        self._shapes = [tuple(a.shape[1:]) for a in arrays]          # per image (heights / widths may differ)
        # a dataset that does not fit beside the activations streams through on its own: without an explicit budget, more than
        # `auto_resident_frac` of the free HBM switches the least-recently-used eviction on (pssr/data.py:553-564 `preload=False`)
        total_bytes = sum(self._sheet_bytes(i) for i in range(len(arrays)))
        if self.max_resident_bytes is None and total_bytes > 0:
            free = torch.cuda.mem_get_info(self.device)[0]
            if total_bytes > self.auto_resident_frac * free:
                self.max_resident_bytes = int(self.auto_resident_frac * free)
        from . import dist as D
        # preload: every upload is queued now, back to back on the side stream (`preload_bytes` bounds it; the rest then follow
        # ahead of the batches).  The bulk copies own the host->device copy engine for a while, which is why a batch's tile table
        # does NOT travel by cudaMemcpyAsync (ops.TileTable / pssr_table_fetch): measured, the second batch of a 50-stack dataset
        # otherwise waits 22 ms behind the last stack.
        if preload and not D.is_dist():
            total = 0
            for i in range(len(arrays)):
                total += self._sheet_bytes(i)
                if i > 0 and ((self.preload_bytes is not None and total > self.preload_bytes) or
                              (self.max_resident_bytes is not None and total > self.max_resident_bytes)):
                    break
                self._ensure(i)

    def _sheet_bytes(self, i):
        f, (h, w) = self._frames_total[i], self._shapes[i]
        return f * h * w * (1 if str(self._sources[i].dtype).endswith("uint8") else 2)

    def _ensure(self, i):
        """Starts the upload of sheet i on the side stream unless it is resident; returns immediately."""
        if self._sheets[i] is not None:
            return
        src = self._sources[i]
        if hasattr(src, "read_pinned"):                  # lazy file source (pssr2_b200/io.py): decoded into pinned staging
            if i + 1 < len(self._sources) and hasattr(self._sources[i + 1], "prefetch"):
                self._sources[i + 1].prefetch()          # the next sheet decodes on the reader thread meanwhile
            src = src.read_pinned()
        t = src if isinstance(src, torch.Tensor) else torch.as_tensor(src.view(np.int16) if src.dtype == np.uint16 else src)
        up = self._upload_stream
        with torch.cuda.device(self.device):
            cur = torch.cuda.current_stream(self.device)
            up.wait_stream(cur)     # the destination block may be recycled memory still in use by kernels queued on `cur`
            # destination from the CURRENT stream's pool: a block allocated under a fresh side stream can never be served from
            # the allocator's cache, i.e. every sheet would cost a synchronous cudaMalloc (measured 0.4 ms per 33 MB sheet)
            d = torch.empty(t.shape, dtype=t.dtype, device=self.device)
            with torch.cuda.stream(up):
                d.copy_(t, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(up)
            d.record_stream(up)
        self._sheets[i] = d
        self._sheet_events[i] = ev
        self._lru.append(i)

    def _evict(self, keep):
        if self.max_resident_bytes is None:
            return
        total = sum(self._sheet_bytes(i) for i in self._lru)
        for i in list(self._lru):
            if total <= self.max_resident_bytes:
                break
            if i in keep:
                continue
            self._lru.remove(i)
            self._sheet_events.pop(i, None)
            self._sheets[i] = None      # the caching allocator frees the block once the kernels queued on it are done
            total -= self._sheet_bytes(i)

    def _wait_sheets(self, sheet_ids):
        """Makes the given sheets resident and orders the current stream after their upload (each event is waited for once)."""
        ids = sorted(set(int(i) for i in sheet_ids))
        for i in ids:
            self._ensure(i)
            if i in self._lru:
                self._lru.remove(i)
                self._lru.append(i)
        cur = torch.cuda.current_stream(self.device)
        for i in ids:
            ev = self._sheet_events.pop(i, None)
            if ev is not None:
                cur.wait_event(ev)
        return ids

    def _prefetch_after(self, ids):
        """Starts the upload of the sheets the NEXT batch will touch: as many as this batch used, following its last one (a batch
        of an ImageDataset spans one stack per item).  Called after the batch's tile table has been queued."""
        keep = set(ids)
        for nxt in range(ids[-1] + 1, min(ids[-1] + 1 + len(ids), len(self._sheets))):
            if self.max_resident_bytes is not None and sum(self._sheet_bytes(i) for i in keep | {nxt}) > self.max_resident_bytes:
                break
            src = self._sources[nxt]
            if not hasattr(src, "read_pinned") or src.ready():
                self._ensure(nxt)            # upload ahead (a file source only once its decode has finished: never block here)
                keep.add(nxt)
            else:
                src.prefetch()
                break
        self._evict(keep)

    def sheet_item_counts(self):
        """Validation items per source image, in ``val_idx`` order (the unit of sheet-aligned sharding): a list of
        (image index, count) runs, or None when the validation items of an image are not contiguous."""
        runs = []
        tiles = getattr(self, "tiles", None)
        for v in self.val_idx:
            img = _get_image_idx(v, self.slices, tiles)[0]
            if runs and runs[-1][0] == img:
                runs[-1][1] += 1
            elif any(r[0] == img for r in runs):
                return None
            else:
                runs.append([img, 1])
        return [(i, c) for i, c in runs]

    # per-item geometry, implemented by subclasses: (sheet, frame0, y, x, vh, vw)
    def _locate(self, idx):
        raise NotImplementedError

    def _frames_window(self, image_idx):
        return max(self.n_frames) if self.n_frames is not None else self._frames_total[image_idx]

    def _table(self, indices, xf=None):
        locs = [self._locate(i) for i in indices]
        cols = list(zip(*locs))
        ids = self._wait_sheets(cols[0])
        table = ops.TileTable(self._sheets, *cols, tile_xf=xf)
        self._prefetch_after(ids)
        return table

    def _draw_rotation(self, idx):
        """pssr/data.py:108 / :244: training items (not validation ones) get a random rot90 and a flip of axis 1, 2 or both, drawn
        from Python's ``random`` exactly as the reference draws them; encoded for the kernel (bit 0 rot90, 1 flip rows, 2 flip
        columns)."""
        if not self.rotation or idx in self._val_set():
            return 0
        rot, axes = bool(random.getrandbits(1)), random.choice((1, 2, (1, 2)))
        axes = (axes,) if isinstance(axes, int) else axes
        return (1 if rot else 0) | (2 if 1 in axes else 0) | (4 if 2 in axes else 0)

    def _val_set(self):
        vs = getattr(self, "_val_set_cache", None)
        if vs is None or vs[0] is not self.val_idx:
            vs = (self.val_idx, set(self.val_idx))
            self._val_set_cache = vs
        return vs[1]

    def loader(self, batch_size, train=True):
        """The data side of ``train_paired`` (pssr/train.py:75-96): yields ``(hr, lr)`` device batches, each from ONE fused
        launch.  ``train=True`` walks the non-validation items in ``random.shuffle`` order with the reference's per-item
        rotation / flip augmentation (``_RandomIterIdx(_invert_idx(val_idx, len))``); ``train=False`` walks the validation items
        in the reference's seeded order (``_RandomIterIdx(val_idx, seed=True)``), unaugmented."""
        if train:
            vs = self._val_set()
            idx = [i for i in range(len(self)) if i not in vs]
            random.shuffle(idx)
        else:
            idx = list(self.val_idx)
            np.random.seed(0)
            np.random.shuffle(idx)
        for s in range(0, len(idx), batch_size):
            b = self.batch(idx[s:s + batch_size], want_hr=not self.is_lr, augment=train)
            yield (b["hr"], b["lr"]) if not self.is_lr else b["lr"]

    def batch(self, indices, want_hr=True, want_hr_u8=False, want_lr=True, tile_index0=None, seed=None, augment=False):
        """One fused launch for many items.  Returns dict(lr=[n,f_lr,h,w] f32, hr=[n,f_hr,H,W] f32 | None,
        hr_u8=[n,1,H,W] u8 | None), all on the device.  (pssr/data.py:100-120 / :236-256 + _gen_pair :471-495)
        ``augment=True``: non-validation items are rotated / flipped at random like the reference's training path."""
        indices = list(indices)
        tiles_ = getattr(self, "tiles", None)
        windows = {self._frames_window(_get_image_idx(i, self.slices, tiles_)[0]) for i in indices}
        crap_ = self.crappifier
        per_item = len(indices) > 1 and (len(windows) > 1 or (isinstance(crap_, Crappifier) and crap_.has_spread()))
        if per_item:
            # items whose frame windows differ (n_frames=-1 over stacks of different depth) cannot share a launch, and a
            # crappifier with spread > 0 redraws its intensity for EVERY item (pssr/crappifiers.py:63,85,104): one launch each
            t0 = indices[0] if tile_index0 is None else tile_index0
            parts = [self.batch([i], want_hr, want_hr_u8, want_lr, t0 + k, seed, augment) for k, i in enumerate(indices)]
            if len(windows) > 1 and not self.is_lr and any(p["lr"].shape != parts[0]["lr"].shape for p in parts):
                raise ValueError("the items of one batch have different frame counts (n_frames=-1 over stacks of different depth); "
                                 "use batch_size=1 or a fixed n_frames")
            return {k: (torch.cat([p[k] for p in parts]) if parts[0][k] is not None else None) for k in ("lr", "hr", "hr_u8")}
        frames = self._frames_window(_get_image_idx(indices[0], self.slices, tiles_)[0])
        xf = getattr(self, "_forced_xf", None)          # paired datasets hand the SAME codes to their HR and LR halves
        if xf is None:
            xf = [self._draw_rotation(i) for i in indices] if augment and not self.is_lr else None
        table = self._table(indices, xf)
        lr_res_scale = self.lr_scale
        if self.is_lr:
            # LR mode (_ready_lr, data.py:518-524): crop/pad only -- identity resample, no noise
            lr, _, _ = ops.crappify(table, self._lr_mode_res, 1, None, frames=frames)
            return {"lr": lr, "hr": None, "hr_u8": None}
        lr0, lrn, hr0, hrn = 0, frames, 0, frames
        if self.n_frames is not None and self.n_frames[0] != self.n_frames[1]:
            if not self.n_frames[1] > frames:
                hr0, hrn = _slice_center_range(frames, self.n_frames[1])
            if not self.n_frames[0] > frames:
                lr0, lrn = _slice_center_range(frames, self.n_frames[0])
        crap = self.crappifier
        specs, clip_between, host_crap = None, False, None
        if crap is not None:
            specs = crap.noise_specs() if isinstance(crap, Crappifier) and hasattr(crap, "noise_specs") else None
            if specs is None:
                host_crap = crap
            else:
                clip_between = bool(getattr(crap, "clip_between", False))
                if len(specs) > 4:
                    host_crap, specs = crap, None      # longer chains than the kernel's four stages run on the host path
        seed = _fresh_seed() if seed is None else seed
        t0 = indices[0] if tile_index0 is None else tile_index0
        if host_crap is None:
            lr, hr, hr8 = ops.crappify(table, self.hr_res, lr_res_scale, specs, frames=frames, lr_frame0=lr0, lr_frames=lrn,
                                       hr_frame0=hr0, hr_frames=hrn, clip_between=clip_between, seed=seed, tile_index0=t0,
                                       want_lr=want_lr, want_hr_f32=want_hr, want_hr_u8=want_hr_u8)
        else:
            # user-supplied callable / host-only Crappifier (data.py:485-486): the device does tile gather + downscale,
            # the user's own function runs on the host, round/clip follows (data.py:487)
            lr_all, hr, hr8 = ops.crappify(table, self.hr_res, lr_res_scale, None, frames=frames, hr_frame0=hr0, hr_frames=hrn,
                                           want_hr_f32=want_hr, want_hr_u8=want_hr_u8)
            res = []
            for tile in lr_all.cpu().numpy():
                out = host_crap.crappify(tile) if isinstance(host_crap, Crappifier) else host_crap(tile)
                out = np.asarray(out.detach().cpu() if isinstance(out, torch.Tensor) else out)
                res.append(np.clip(out.round(), 0, 255)[lr0:lr0 + lrn].astype(np.float32))
            lr = torch.as_tensor(np.stack(res)).to(self.device)
        return {"lr": lr, "hr": hr, "hr_u8": hr8}

    def __getitem__(self, idx, pp=False):
        if idx >= len(self):
            raise IndexError(f"Tried to retrieve invalid image. Index {idx} is not less than {len(self)} total image frame slices.")
        b = self.batch([idx], augment=not pp)      # validation items and pp=True are never augmented (data.py:103,108)
        if self.is_lr:
            return b["lr"][0]
        return b["hr"][0], b["lr"][0]


class SlidingDataset(_DeviceDataset):
    def __init__(self, path, hr_res: int = 512, lr_scale: int = 4, crappifier=Poisson(), overlap: int = 128, n_frames=-1,
                 slide: bool = False, stack: str = "TZ", extension: str = "czi", preload: bool = True, val_split: float = 0.1,
                 rotation: bool = True, split_seed: int = 0, extra_path=None, extra_scale: int = 1, transforms=None,
                 device="cuda"):
        r"""Tiles image sheets into overlapping ``hr_res`` tiles and returns crappified high/low-resolution
        pairs (pssr/data.py:132-266).  Sheets stay resident on ``device``; see the module docstring."""
        super().__init__()
        if extra_path is not None or transforms is not None:
            raise NotImplementedError("extra_path / transforms are training-time options outside the accelerated predict path")
        self.path = path
        names, arrays = _load_sources(path, extension, stack=stack.upper())
        self.hr_files = names
        overlap = 0 if overlap is None else overlap
        if not hr_res > overlap:
            raise ValueError(f"hr_res must be greater than overlap. Given values are {hr_res} and {overlap} respectively.")
        self.stride = hr_res - overlap
        self.stack = stack.upper()
        lr_scale = None if lr_scale == -1 else lr_scale
        self.n_frames = _get_n_frames(n_frames)
        self.slide = slide
        self.preload = preload
        self.tiles, self.slices, self._tiles_y = [], [], []
        for a in arrays:
            tx, ty = _n_tiles(a.shape[-2:], hr_res, self.stride)
            self.tiles.append(tx * ty)
            self._tiles_y.append(ty)
            self.slices.append(1 if self.n_frames is None else
                               ((a.shape[0] - max(self.n_frames) + 1) if slide else (a.shape[0] // max(self.n_frames))))
        self.val_idx = _get_val_idx(self.slices, val_split, split_seed, self.tiles)
        self.crop_res = hr_res
        self.is_lr = lr_scale is None
        if self.is_lr:
            print("LR mode is enabled, dataset will load only unmodified low-resolution images.")
            if val_split < 1:
                warnings.warn("val_split is less than 1, not all low-resolution images will be used in prediciton.", stacklevel=2)
        self.hr_res, self.lr_scale = hr_res, lr_scale
        self._lr_mode_res = hr_res
        self.crappifier, self.rotation, self.extra_scale, self.transforms = crappifier, rotation, extra_scale, transforms
        self._upload(arrays, device, preload=preload)

    def _locate(self, idx):
        image_idx, local = _get_image_idx(idx, self.slices, self.tiles)
        n_slices = self.slices[image_idx]
        tile_idx = local // n_slices
        ty = self._tiles_y[image_idx]
        y, x = tile_idx // ty * self.stride, tile_idx % ty * self.stride          # data.py:633-634
        f0 = 0 if self.n_frames is None else (local % n_slices) * (1 if self.slide else max(self.n_frames))
        return image_idx, f0, y, x, self.hr_res, self.hr_res

    def __len__(self):
        return sum(t * s for t, s in zip(self.tiles, self.slices))

    def __repr__(self):
        res = f"low-res: {self.hr_res}" if self.is_lr else f"high-res: {self.hr_res}, low-res: {self.hr_res // self.lr_scale}"
        return f'SlidingDataset from path "{self.path if isinstance(self.path, (str, Path)) else "<memory>"}"\n{len(self.hr_files)} files with {len(self)} total frame slices\n{res}'

    def _get_name(self, idx):
        image_idx, idx = _get_image_idx(idx, self.slices, self.tiles)
        return f"{self.hr_files[image_idx].split('.')[0]}_{idx // self.slices[image_idx]}_{idx % self.slices[image_idx]}"


class ImageDataset(_DeviceDataset):
    def __init__(self, path, hr_res: int = 512, lr_scale: int = 4, crappifier=Poisson(), n_frames=-1, extension: str = "tif",
                 val_split: float = 0.1, rotation: bool = True, split_seed: int = 0, extra_path=None, extra_scale: int = 1,
                 transforms=None, device="cuda"):
        r"""Pre-tiled images -> centre crop / reflect pad to ``hr_res`` -> crappified pairs (pssr/data.py:12-130)."""
        super().__init__()
        if extra_path is not None or transforms is not None:
            raise NotImplementedError("extra_path / transforms are training-time options outside the accelerated predict path")
        self.path = path
        names, arrays = _load_sources(path, extension)
        self.hr_files = names
        lr_scale = None if lr_scale == -1 else lr_scale
        self.n_frames = _get_n_frames(n_frames)
        self.slices, max_size = [], 0
        for a in arrays:
            self.slices.append(1 if self.n_frames is None else a.shape[0] // max(self.n_frames))
            max_size = max(max(a.shape[-2:]), max_size)
        self.val_idx = _get_val_idx(self.slices, val_split, split_seed)
        self.crop_res = min(hr_res, max_size)
        self.is_lr = lr_scale is None or max_size <= hr_res // lr_scale
        if self.is_lr:
            print("LR mode is enabled, dataset will load only unmodified low-resolution images.")
            if val_split < 1:
                warnings.warn("val_split is less than 1, not all low-resolution images will be used in prediciton.", stacklevel=2)
        self.hr_res = hr_res
        self.lr_scale = lr_scale if lr_scale is not None else 1
        self._lr_mode_res = hr_res // self.lr_scale
        self.crappifier, self.rotation, self.extra_scale, self.transforms = crappifier, rotation, extra_scale, transforms
        self._upload(arrays, device)

    def _locate(self, idx):
        image_idx, local = _get_image_idx(idx, self.slices)
        h, w = self._shapes[image_idx]
        res = self._lr_mode_res if self.is_lr else self.hr_res
        if [h, w] == [res] * 2:                                   # _square_crop, data.py:536-546
            y = x = 0
            size = res
        else:
            size = min(h, w, res)
            y, x = (h - size) // 2, (w - size) // 2
        f0 = 0 if self.n_frames is None else (local % self.slices[image_idx]) * max(self.n_frames)
        return image_idx, f0, y, x, size, size

    def __len__(self):
        return sum(self.slices)

    def __repr__(self):
        res = f"low-res: {self.hr_res // self.lr_scale}" if self.is_lr else f"high-res: {self.hr_res}, low-res: {self.hr_res // self.lr_scale}"
        return f'ImageDataset from path "{self.path if isinstance(self.path, (str, Path)) else "<memory>"}"\n{len(self.hr_files)} files with {len(self)} total frame slices\n{res}'

    def _get_name(self, idx):
        image_idx, idx = _get_image_idx(idx, self.slices)
        return self.hr_files[image_idx].split('.')[0] + (f"_{idx}" if self.n_frames is not None else "")


class _PairedDataset(Dataset):
    """Ground-truth high- / low-resolution pairs without a crappifier (pssr/data.py:268-346, :348-470, `_transform_pair` :497-516).
    Two device-resident halves -- the HR images at ``hr_res`` and the LR images at ``hr_res // lr_scale``, each cropped / padded by
    the gather kernel in its identity-resample mode -- indexed together and augmented with the same rot90 / flip code."""

    is_lr = False
    crappifier = None
    extra_hr_files = None

    def _pair(self, hr_ds, lr_ds, hr_res, lr_scale, rotation, nf, val_split, split_seed):
        self._hr, self._lr = hr_ds, lr_ds
        if len(hr_ds.hr_files) != len(lr_ds.hr_files):
            raise FileNotFoundError(f"Mismatch between amounts of high-low-resolution images. Found {len(hr_ds.hr_files)} high-resolution "
                                    f"and {len(lr_ds.hr_files)} low-resolution images.")
        self.hr_files, self.lr_files = hr_ds.hr_files, lr_ds.hr_files
        # the reference counts slices from the HR stack and max(n_frames) (data.py:312) and reads slice i of BOTH stacks; the halves
        # count with their own frame window, so an item is addressed as (image, slice) in each of them
        nfm = None if nf is None else max(nf)
        self.slices = [1 if nfm is None else f // nfm for f in hr_ds._frames_total]
        tiles = getattr(hr_ds, "tiles", None)
        self.val_idx = _get_val_idx(self.slices, val_split, split_seed, tiles)
        self.n_frames = nf
        self.hr_res, self.lr_scale, self.rotation = hr_res, lr_scale, rotation
        self.crop_res = hr_ds.crop_res
        self.device = hr_ds.device

    def __len__(self):
        tiles = getattr(self, "tiles", None)
        return sum(s * (t if tiles else 1) for s, t in zip(self.slices, tiles or [1] * len(self.slices)))

    def _inner(self, ds, idx):
        """Flat index of item ``idx`` (counted with this dataset's slices) inside one half (counted with its own)."""
        tiles = getattr(self, "tiles", None)
        image_idx, local = _get_image_idx(idx, self.slices, tiles)
        if tiles:
            tile, sl = local // self.slices[image_idx], local % self.slices[image_idx]
            local = tile * ds.slices[image_idx] + sl
        return sum(s * (t if tiles else 1) for s, t in zip(ds.slices[:image_idx], (tiles or [1] * len(ds.slices))[:image_idx])) + local

    def batch(self, indices, augment=False, **_):
        indices = list(indices)
        xf = None
        if augment and self.rotation:
            vs = set(self.val_idx)
            xf = []
            for i in indices:
                if i in vs:
                    xf.append(0)
                else:
                    rot, axes = bool(random.getrandbits(1)), random.choice((1, 2, (1, 2)))
                    axes = (axes,) if isinstance(axes, int) else axes
                    xf.append((1 if rot else 0) | (2 if 1 in axes else 0) | (4 if 2 in axes else 0))
        self._hr._forced_xf = self._lr._forced_xf = xf
        try:
            hr = self._hr.batch([self._inner(self._hr, i) for i in indices])["lr"]      # identity-resample mode: the cropped / padded tile
            lr = self._lr.batch([self._inner(self._lr, i) for i in indices])["lr"]
        finally:
            self._hr._forced_xf = self._lr._forced_xf = None
        return {"hr": hr, "lr": lr, "hr_u8": None}

    def __getitem__(self, idx, pp=False):
        if idx >= len(self):
            raise IndexError(f"Tried to retrieve invalid image. Index {idx} is not less than {len(self)} total image frame slices.")
        b = self.batch([idx], augment=not pp)
        return b["hr"][0], b["lr"][0]

    def loader(self, batch_size, train=True):
        """See ``_DeviceDataset.loader``: paired batches for ``train_paired`` (pssr/train.py:75-96)."""
        if train:
            vs = set(self.val_idx)
            idx = [i for i in range(len(self)) if i not in vs]
            random.shuffle(idx)
        else:
            idx = list(self.val_idx)
            np.random.seed(0)
            np.random.shuffle(idx)
        for s in range(0, len(idx), batch_size):
            b = self.batch(idx[s:s + batch_size], augment=train)
            yield b["hr"], b["lr"]


def _silence(fn, *a, **k):
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return fn(*a, **k)


class PairedImageDataset(_PairedDataset):
    def __init__(self, hr_path, lr_path, hr_res: int = 512, lr_scale: int = 4, n_frames=-1, extension: str = "tif", val_split: float = 1,
                 rotation: bool = True, split_seed: int = None, transforms=None, device="cuda"):
        r"""Paired pre-tiled high- / low-resolution images (pssr/data.py:268-346)."""
        super().__init__()
        if transforms is not None:
            raise NotImplementedError("transforms are outside the accelerated path")
        nf = _get_n_frames(n_frames)
        hr = _silence(ImageDataset, hr_path, hr_res=hr_res, lr_scale=-1, crappifier=None, n_frames=-1 if nf is None else nf[1], extension=extension,
                      val_split=val_split, rotation=False, split_seed=split_seed, device=device)
        lr = _silence(ImageDataset, lr_path, hr_res=hr_res // lr_scale, lr_scale=-1, crappifier=None, n_frames=-1 if nf is None else nf[0],
                      extension=extension, val_split=val_split, rotation=False, split_seed=split_seed, device=device)
        self._pair(hr, lr, hr_res, lr_scale, rotation, nf, val_split, split_seed)

    def _get_name(self, idx):
        image_idx, idx = _get_image_idx(idx, self.slices)
        return self.lr_files[image_idx].split('.')[0] + (f"_{idx}" if self.n_frames is not None else "")


class PairedSlidingDataset(_PairedDataset):
    def __init__(self, hr_path, lr_path, hr_res: int = 512, lr_scale: int = 4, overlap: int = 128, n_frames=-1, slide: bool = False,
                 stack: str = "TZ", extension: str = "tif", preload: bool = True, val_split: float = 1, rotation: bool = True,
                 split_seed: int = None, transforms=None, device="cuda"):
        r"""Paired high- / low-resolution image sheets tiled with matching windows (pssr/data.py:348-470)."""
        super().__init__()
        if transforms is not None:
            raise NotImplementedError("transforms are outside the accelerated path")
        nf = _get_n_frames(n_frames)
        hr = _silence(SlidingDataset, hr_path, hr_res=hr_res, lr_scale=-1, crappifier=None, overlap=overlap, n_frames=-1 if nf is None else nf[1],
                      slide=slide, stack=stack, extension=extension, preload=preload, val_split=val_split, rotation=False,
                      split_seed=split_seed, device=device)
        lr = _silence(SlidingDataset, lr_path, hr_res=hr_res // lr_scale, lr_scale=-1, crappifier=None, overlap=(overlap or 0) // lr_scale,
                      n_frames=-1 if nf is None else nf[0], slide=slide, stack=stack, extension=extension, preload=preload,
                      val_split=val_split, rotation=False, split_seed=split_seed, device=device)
        self.tiles = hr.tiles
        self._pair(hr, lr, hr_res, lr_scale, rotation, nf, val_split, split_seed)

    def _get_name(self, idx):
        image_idx, idx = _get_image_idx(idx, self.slices, self.tiles)
        return f"{self.lr_files[image_idx].split('.')[0]}_{idx // self.slices[image_idx]}_{idx % self.slices[image_idx]}"
