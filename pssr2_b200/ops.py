"""Thin torch-tensor wrappers over the C ABI for kernel families 1, 3 and 4
(``include/pssr_b200.h``).  torch provides device memory and the current stream only.
"""
import ctypes
import math
import threading

import numpy as np
import torch

from . import _lib
from ._lib import (NOISE_GAUSSIAN, NOISE_POISSON, NOISE_SALTPEPPER, RNG_INJECTED, RNG_PHILOX, CrappifyArgs, NoiseStage)


def _cuda(t, what):
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise RuntimeError(f"{what} must be a CUDA tensor (pssr2_b200 has no CPU path)")
    return t


class NoiseSpec:
    """One resolved noise stage of a crappify call."""

    def __init__(self, kind, intensity=0.0, gain=0.0, mix_in_f32=True, injected=None):
        self.kind, self.intensity, self.gain, self.mix_in_f32, self.injected = kind, float(intensity), float(gain), mix_in_f32, injected


class _TableStaging:
    """Pinned staging buffers of the tile tables.  The table reaches the device through a kernel that reads the pinned buffer in
    place (``pssr_table_fetch``), which torch's pinned-memory allocator knows nothing about: a buffer is handed out again only
    after the event recorded behind its fetch kernel has completed."""

    def __init__(self):
        self._idle = {}                  # device -> [(pinned uint8 tensor, event recorded after its last fetch)]
        self._lock = threading.Lock()

    def take(self, nbytes, dev):
        with self._lock:
            pool = self._idle.setdefault(dev, [])
            for k, (buf, ev) in enumerate(pool):
                if buf.numel() >= nbytes and ev.query():
                    pool.pop(k)
                    return buf
        return torch.empty(max(int(nbytes), 4096), dtype=torch.uint8, pin_memory=True)

    def give_back(self, buf, dev):
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(dev))
        with self._lock:
            self._idle.setdefault(dev, []).append((buf, ev))


_table_staging = _TableStaging()


class TileTable:
    """Device-resident description of where each HR tile lives inside the resident sheets.

    ``sheets`` may hold ``None`` for sheets that are not resident; only the sheets the tiles reference are used (their indices
    are remapped), and the pointer table, the per-sheet dimensions and the six index columns travel in ONE pinned-host ->
    device copy on the current stream."""

    def __init__(self, sheets, tile_sheet, tile_frame, tile_y, tile_x, tile_vh, tile_vw, tile_xf=None):
        """tile_xf: optional per-tile augmentation code (bit 0 rot90, bit 1 flip rows, bit 2 flip columns; pssr/data.py:478-480)."""
        tile_sheet = np.asarray(tile_sheet, dtype=np.int64)
        used = sorted(set(int(i) for i in tile_sheet))
        remap = {g: k for k, g in enumerate(used)}
        self.sheets = [_cuda(sheets[g], "sheet").contiguous() for g in used]
        dev = _lib.same_device(*self.sheets)
        s0 = self.sheets[0]
        for s in self.sheets:
            if s.dim() != 3 or s.dtype != s0.dtype:
                raise ValueError("all sheets must be [frames, H, W] with one dtype")
        if s0.dtype == torch.uint8:
            self.elem_bytes = 1
        elif s0.dtype in (torch.uint16, torch.int16):
            self.elem_bytes = 2
        else:
            raise TypeError(f"sheets must be uint8 or uint16, got {s0.dtype}")
        self.sheet_h, self.sheet_w = int(s0.shape[1]), int(s0.shape[2])
        n, ns = int(tile_sheet.size), len(used)
        local = np.asarray([remap[int(i)] for i in tile_sheet], dtype=np.int32)
        self.frames_of_tile = np.asarray([int(self.sheets[k].shape[0]) for k in local], dtype=np.int64)   # host copy: frame-range checks
        self.tile_frame_host = np.asarray(tile_frame, dtype=np.int64)
        # packed header: [ns] int64 pointers | [2*ns] int32 sheet dims | [7*n] int32 columns (the seventh: augmentation codes)
        nbytes = 8 * ns + 4 * (2 * ns + 7 * n)
        host = _table_staging.take(nbytes, dev)
        hv = host.numpy()[:nbytes]
        hv[:8 * ns].view(np.int64)[:] = [s.data_ptr() for s in self.sheets]
        i32 = hv[8 * ns:].view(np.int32)
        i32[:ns] = [int(s.shape[1]) for s in self.sheets]
        i32[ns:2 * ns] = [int(s.shape[2]) for s in self.sheets]
        xf = np.zeros(n, dtype=np.int32) if tile_xf is None else np.asarray(tile_xf, dtype=np.int32)
        i32[2 * ns:] = np.asarray([local, tile_frame, tile_y, tile_x, tile_vh, tile_vw, xf], dtype=np.int32).reshape(-1)
        # a one-CTA kernel reads the pinned buffer in place (pssr_table_fetch): a copy-engine transfer would queue behind the bulk
        # sheet uploads already handed to the host->device engine
        packed = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        with _lib.on_device(dev):
            _lib.check(_lib.lib().pssr_table_fetch(packed.data_ptr(), host.data_ptr(), nbytes, _lib.current_stream_ptr(dev)))
            _table_staging.give_back(host, dev)            # reusable once the stream has passed the fetch kernel
        self._packed = packed
        self.ptrs = packed[:8 * ns].view(torch.int64)
        d32 = packed[8 * ns:].view(torch.int32)
        # sheets of different sizes: per-sheet dimension arrays travel with the table
        self.sheet_dims = d32[:2 * ns].view(2, ns) if any(s.shape[1:] != s0.shape[1:] for s in self.sheets) else None
        cols = d32[2 * ns:].view(7, n)
        self.tile_sheet, self.tile_frame, self.tile_y, self.tile_x, self.tile_vh, self.tile_vw = (cols[i] for i in range(6))
        self.tile_xf = cols[6] if tile_xf is not None and bool(xf.any()) else None
        self.n_tiles = n
        self.device = dev

    def slice(self, start, count):
        t = object.__new__(TileTable)
        t.__dict__.update(self.__dict__)
        for k in ("tile_sheet", "tile_frame", "tile_y", "tile_x", "tile_vh", "tile_vw"):
            setattr(t, k, getattr(self, k)[start:start + count])
        t.tile_xf = self.tile_xf[start:start + count] if self.tile_xf is not None else None
        t.frames_of_tile = self.frames_of_tile[start:start + count]
        t.tile_frame_host = self.tile_frame_host[start:start + count]
        t.n_tiles = count
        return t


def crappify(table: TileTable, hr_res, lr_scale, stages, *, frames=1, lr_frame0=0, lr_frames=None, hr_frame0=0, hr_frames=None,
             clip_between=False, seed=0, tile_index0=0, want_lr=True, want_hr_f32=False, want_hr_u8=False):
    """Fused tile gather + Pillow-exact BILINEAR downscale + noise + round/clip (pssr/data.py:471-495).
    stages: list[NoiseSpec] or None (crappifier=None).  Returns (lr f32 [n, lr_frames, lr, lr] | None,
    hr f32 [n, hr_frames, hr, hr] | None, hr u8 [n, 1, hr, hr] | None)."""
    n = table.n_tiles
    lr_res = hr_res // lr_scale
    lr_frames = frames if lr_frames is None else lr_frames
    hr_frames = frames if hr_frames is None else hr_frames
    dev = table.device
    # the kernel reads frames [tile_frame, tile_frame + frames) of each tile's sheet: a window that runs past a (shorter) sheet
    # would be an out-of-bounds device read, so it is rejected here
    bad = np.nonzero((table.tile_frame_host < 0) | (table.tile_frame_host + frames > table.frames_of_tile))[0]
    if bad.size:
        k = int(bad[0])
        raise ValueError(f"tile {k}: frames [{int(table.tile_frame_host[k])}, {int(table.tile_frame_host[k]) + frames}) exceed its sheet's "
                         f"{int(table.frames_of_tile[k])} frames (one batch must not mix sheets with different frame windows)")
    a = CrappifyArgs()
    a.sheets = table.ptrs.data_ptr()
    a.n_sheets = len(table.sheets)
    a.elem_bytes = table.elem_bytes
    a.sheet_h, a.sheet_w = table.sheet_h, table.sheet_w
    if table.sheet_dims is not None:
        a.sheet_hs, a.sheet_ws = table.sheet_dims[0].data_ptr(), table.sheet_dims[1].data_ptr()
    a.tile_sheet, a.tile_frame = table.tile_sheet.data_ptr(), table.tile_frame.data_ptr()
    a.tile_y, a.tile_x = table.tile_y.data_ptr(), table.tile_x.data_ptr()
    a.tile_vh, a.tile_vw = table.tile_vh.data_ptr(), table.tile_vw.data_ptr()
    a.tile_xf = table.tile_xf.data_ptr() if table.tile_xf is not None else None
    a.n_tiles, a.frames, a.hr_res, a.lr_scale = n, frames, hr_res, lr_scale
    keep = []
    stages = [] if stages is None else stages
    if len(stages) > 4:
        raise ValueError("at most 4 chained noise stages are supported")
    for i, s in enumerate(stages):
        st = NoiseStage()
        st.kind = s.kind
        st.intensity, st.gain = s.intensity, s.gain
        st.mix_in_f32 = 1 if s.mix_in_f32 else 0
        if s.injected is not None:
            inj = _cuda(s.injected, "injected noise").contiguous()
            want = {NOISE_POISSON: torch.int64, NOISE_GAUSSIAN: torch.float64, NOISE_SALTPEPPER: torch.uint8}[s.kind]
            if inj.dtype != want or inj.numel() != n * frames * lr_res * lr_res:
                raise ValueError(f"injected noise for stage {i} must be {want} with {n * frames * lr_res * lr_res} elements")
            keep.append(inj)
            st.rng, st.injected = RNG_INJECTED, inj.data_ptr()
        else:
            st.rng, st.injected = RNG_PHILOX, None
        a.stages[i] = st
    a.n_stages = len(stages)
    a.clip_between = 1 if clip_between else 0
    a.seed, a.tile_index0 = int(seed) & (2 ** 64 - 1), int(tile_index0)
    a.hr_frame0, a.hr_frames, a.lr_frame0, a.lr_frames = hr_frame0, hr_frames, lr_frame0, lr_frames
    lr = torch.empty(n, lr_frames, lr_res, lr_res, dtype=torch.float32, device=dev) if want_lr else None
    hr = torch.empty(n, hr_frames, hr_res, hr_res, dtype=torch.float32, device=dev) if want_hr_f32 else None
    hr8 = torch.empty(n, 1, hr_res, hr_res, dtype=torch.uint8, device=dev) if want_hr_u8 else None
    a.lr_out = lr.data_ptr() if lr is not None else None
    a.hr_out = hr.data_ptr() if hr is not None else None
    a.hr_u8_out = hr8.data_ptr() if hr8 is not None else None
    with _lib.on_device(dev):
        _lib.check(_lib.lib().pssr_crappify(ctypes.byref(a), _lib.current_stream_ptr(dev)), "pssr_crappify")
    return lr, hr, hr8


def resize_bilinear(img: torch.Tensor, scale: int) -> torch.Tensor:
    """Pillow-exact `Image.resize(BILINEAR)` by an integer factor on [n, h, w] uint8/uint16."""
    img = _cuda(img, "img").contiguous()
    n, h, w = img.shape
    eb = img.element_size()
    out = torch.empty(n, h // scale, w // scale, dtype=img.dtype, device=img.device)
    with _lib.on_device(img.device):
        _lib.check(_lib.lib().pssr_resize_bilinear(img.data_ptr(), out.data_ptr(), n, h, w, scale, eb, _lib.current_stream_ptr(img.device)),
                   "pssr_resize_bilinear")
    return out


def stitch(tiles: torch.Tensor, n_rows, n_cols, overlap, margin) -> torch.Tensor:
    """`_patch_images` + uint8 cast (pssr/util.py:96-137): tiles [stacks*n_rows*n_cols, T, T] uint8."""
    tiles = _cuda(tiles, "tiles").contiguous()
    if tiles.dtype != torch.uint8 or tiles.dim() != 3 or tiles.shape[1] != tiles.shape[2]:
        raise ValueError("tiles must be uint8 [n, T, T]")
    per = n_rows * n_cols
    stacks = tiles.shape[0] // per
    T = int(tiles.shape[1])
    step = T - overlap
    out = torch.empty(stacks, n_rows * step + overlap, n_cols * step + overlap, dtype=torch.uint8, device=tiles.device)
    with _lib.on_device(tiles.device):
        rc = _lib.lib().pssr_stitch(tiles.data_ptr(), out.data_ptr(), stacks, n_rows, n_cols, T, overlap, margin,
                                    _lib.current_stream_ptr(tiles.device))
    if rc == -1 and b"margin" in _lib.lib().pssr_last_error():
        raise ValueError(_lib.lib().pssr_last_error().decode())
    _lib.check(rc, "pssr_stitch")
    return out


def metric_sums(a: torch.Tensor, b: torch.Tensor, want_ssim=True):
    """Exact per-image sums for uint8 pairs [n, h, w]: (sum (a-b)^2 int64 [n], SSIM-map sum float64 [n] | None)."""
    a, b = _cuda(a, "a").contiguous(), _cuda(b, "b").contiguous()
    _lib.same_device(a, b)
    if a.dtype != torch.uint8 or b.dtype != torch.uint8 or a.shape != b.shape or a.dim() != 3:
        raise ValueError("metric_sums expects two uint8 tensors of identical shape [n, h, w]")
    n, h, w = a.shape
    sq = torch.empty(n, dtype=torch.int64, device=a.device)
    ss = torch.empty(n, dtype=torch.float64, device=a.device) if want_ssim else None
    ws = torch.empty(max(16, int(_lib.lib().pssr_metric_workspace_bytes(n, h, w))), dtype=torch.uint8, device=a.device)
    with _lib.on_device(a.device):
        rc = _lib.lib().pssr_metric_sums(a.data_ptr(), b.data_ptr(), n, h, w, sq.data_ptr(), ss.data_ptr() if ss is not None else None,
                                         ws.data_ptr(), _lib.current_stream_ptr(a.device))
    if rc == -1 and b"win_size" in _lib.lib().pssr_last_error():
        raise ValueError(_lib.lib().pssr_last_error().decode())
    _lib.check(rc, "pssr_metric_sums")
    return sq, ss


def normalize_preds_u8(hr: torch.Tensor, hr_hat: torch.Tensor, pmin=0.1, pmax=99.9):
    """`normalize_preds` (pssr/util.py:139-191) on device for uint8 pairs [n, h, w] of equal shape."""
    hr, hr_hat = _cuda(hr, "hr").contiguous(), _cuda(hr_hat, "hr_hat").contiguous()
    _lib.same_device(hr, hr_hat)
    if hr.dtype != torch.uint8 or hr_hat.dtype != torch.uint8 or hr.shape != hr_hat.shape or hr.dim() != 3:
        raise ValueError("normalize_preds_u8 expects two uint8 tensors of identical shape [n, h, w]")
    n, h, w = hr.shape
    ws = torch.empty(int(_lib.lib().pssr_normalize_workspace_bytes(n)), dtype=torch.uint8, device=hr.device)
    oa, ob = torch.empty_like(hr), torch.empty_like(hr_hat)
    with _lib.on_device(hr.device):
        _lib.check(_lib.lib().pssr_normalize_preds(hr.data_ptr(), hr_hat.data_ptr(), oa.data_ptr(), ob.data_ptr(), n, h, w, pmin, pmax,
                                                   ws.data_ptr(), _lib.current_stream_ptr(hr.device)), "pssr_normalize_preds")
    return oa, ob


def normalize_preds_resized_u8(hr: torch.Tensor, hr_hat: torch.Tensor, pmin=0.1, pmax=99.9):
    """normalize_preds (pssr/util.py:139-191) for uint8 hr [n,H,W] and a LOWER-resolution hr_hat [n,h,w] (util.py:179)."""
    hr, hr_hat = _cuda(hr, "hr").contiguous(), _cuda(hr_hat, "hr_hat").contiguous()
    _lib.same_device(hr, hr_hat)
    n, h, w = hr.shape
    hh, hw = hr_hat.shape[-2:]
    oa, ob = torch.empty_like(hr), torch.empty_like(hr_hat)
    with _lib.on_device(hr.device):
        ws = torch.empty(int(_lib.lib().pssr_normalize_resized_workspace_bytes(n)), dtype=torch.uint8, device=hr.device)
        _lib.check(_lib.lib().pssr_normalize_preds_resized(hr.data_ptr(), hr_hat.data_ptr(), oa.data_ptr(), ob.data_ptr(), n, h, w, hh, hw,
                                                           pmin, pmax, ws.data_ptr(), _lib.current_stream_ptr(hr.device)),
                   "pssr_normalize_preds_resized")
    return oa, ob


def profile_hist(a: torch.Tensor, base: torch.Tensor):
    """Noise-profile histogram of `approximate_crappifier`'s objective (pssr/train.py:366-380): (int64 [511] counts of
    float32(a) - float32(base) over np.arange(-256, 256), float64 sum of the profile).  a: uint8 / float32 / float64, base: uint8."""
    a, base = _cuda(a, "a").contiguous(), _cuda(base, "base").contiguous()
    _lib.same_device(a, base)
    kind = {torch.uint8: 0, torch.float32: 1, torch.float64: 2}.get(a.dtype)
    if kind is None or base.dtype != torch.uint8 or a.numel() != base.numel():
        raise ValueError("profile_hist expects a uint8 / float32 / float64 tensor and a uint8 base of the same size")
    hist = torch.empty(511, dtype=torch.int64, device=a.device)
    total = torch.empty(1, dtype=torch.float64, device=a.device)
    with _lib.on_device(a.device):
        _lib.check(_lib.lib().pssr_profile_hist(a.data_ptr(), kind, base.data_ptr(), a.numel(), hist.data_ptr(), total.data_ptr(),
                                                _lib.current_stream_ptr(a.device)), "pssr_profile_hist")
    return hist, total
