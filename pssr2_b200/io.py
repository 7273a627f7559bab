"""File I/O edges of the predict path.

The reference reads sheets with ``tifffile.imread`` / ``czifile`` / Pillow (pssr/data.py:566-627) and writes
predictions and stitched sheets with ``tifffile.imwrite`` (pssr/predict.py:71, pssr/util.py:103), synchronously on
the thread that drives the GPU.  Here

* ``SheetFile`` is a lazy sheet source: its geometry comes from a header probe, its pixels are decoded by the
  library's TIFF reader (``pssr_tiff_read``, ctypes releases the GIL) on a reader thread straight into PINNED host
  memory, so a dataset's side-stream upload starts from the decode buffer without another copy and the decode of
  sheet i+1 overlaps the prediction of sheet i;
* ``write_tiff`` / ``TiffWriter`` encode uint8 / uint16 stacks with ``pssr_tiff_write``; the writer runs on a small
  thread pool so ``predict_images(out_dir=...)`` never waits for the disk;
* ``read_czi`` decodes uncompressed Zeiss CZI (ZISRAW) sub-blocks and ``czi_to_sheet`` restates the reference's
  axis selection, channel mean, stack order, flattening and max-normalisation to uint8 (pssr/data.py:585-619).
  czifile is not installed in this environment: the container layout follows the published ZISRAW structure
  (segment headers, sub-block directory, dimension entries) and is PARITY-UNPINNED against czifile itself.

Compressed / tiled / colour TIFFs and the other Pillow formats fall back to Pillow (``.convert(mode)`` -> uint8,
as pssr/data.py:640-647 does).  Host code only; nothing here touches the device.
"""
import ctypes
import os
import struct
import threading
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

from . import _lib


def _can_pin():
    return torch.cuda.is_available()       # page-locking needs the driver; a CPU-only box (header probes, tests) gets pageable memory


# ------------------------------------------------------------------------------------ TIFF
def tiff_probe(path):
    """-> (frames, h, w, bits, native) from the file header (no pixel is read)."""
    v = [ctypes.c_int32() for _ in range(5)]
    _lib.check(_lib.lib().pssr_tiff_probe(os.fsencode(str(path)), *[ctypes.byref(x) for x in v]), "pssr_tiff_probe")
    return tuple(int(x.value) for x in v[:4]) + (bool(v[4].value),)


def _pillow_stack(path, mode=None):
    """Pillow decode: TIFF keeps its native 8 / 16-bit depth (what tifffile returns), other formats are converted to ``mode``."""
    from PIL import Image
    im = Image.open(path)
    frames = []
    for i in range(getattr(im, "n_frames", 1)):
        im.seek(i)
        fr = im
        if mode is not None:
            fr = fr.convert(mode)                       # pssr/data.py:640-647 `_frame_channel`
        elif fr.mode not in ("L", "I;16", "I;16L", "I;16B", "1", "P"):
            fr = fr.convert("L")
        a = np.asarray(fr)
        if a.dtype == np.bool_:
            a = a.astype(np.uint8) * 255
        if a.dtype.byteorder == ">":
            a = a.astype(a.dtype.newbyteorder("<"))
        frames.append(np.asarray(a, dtype=np.uint8 if mode is not None else a.dtype))
    return np.stack(frames)


def read_tiff(path, pin=False):
    """[frames, H, W] uint8 / uint16 (tifffile.imread + the frame axis of pssr/data.py:569-571).  ``pin=True`` returns a
    pinned torch tensor (int16 container for 16-bit data) the upload can start from directly."""
    frames, h, w, bits, native = tiff_probe(path)
    if not native:
        a = _pillow_stack(path)
        if not pin:
            return a
        t = torch.as_tensor(a.view(np.int16) if a.dtype == np.uint16 else a)
        return t.pin_memory() if _can_pin() else t
    dt = torch.uint8 if bits == 8 else torch.int16
    t = torch.empty(frames, h, w, dtype=dt, pin_memory=pin and _can_pin())
    _lib.check(_lib.lib().pssr_tiff_read(os.fsencode(str(path)), t.data_ptr(), t.numel() * t.element_size()), "pssr_tiff_read")
    if pin:
        return t
    a = t.numpy()
    return a.view(np.uint16) if bits == 16 else a


def write_tiff(path, arr):
    """uint8 / uint16 [H, W] or [frames, H, W] (or [.., 1, H, W]) -> TIFF (tifffile.imwrite's role, pssr/predict.py:71)."""
    a = np.asarray(arr)
    if a.dtype == np.int16:
        a = a.view(np.uint16)
    if a.dtype not in (np.uint8, np.uint16):
        raise TypeError(f"write_tiff expects uint8 or uint16, got {a.dtype}")
    a = np.ascontiguousarray(a.reshape((-1,) + a.shape[-2:]))
    _lib.check(_lib.lib().pssr_tiff_write(os.fsencode(str(path)), a.ctypes.data, a.shape[0], a.shape[1], a.shape[2], a.dtype.itemsize * 8),
               "pssr_tiff_write")


class TiffWriter:
    """Asynchronous ``write_tiff``: encodes on a thread pool (the C call runs without the GIL); ``close()`` waits for the files."""

    def __init__(self, threads=4):
        self._pool = ThreadPoolExecutor(max_workers=threads)
        self._pending = []

    def write(self, path, arr):
        self._pending.append(self._pool.submit(write_tiff, path, arr))

    def close(self):
        for f in self._pending:
            f.result()              # re-raises a failed write
        self._pending = []
        self._pool.shutdown(wait=True)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


# ------------------------------------------------------------------------------------ lazy sheet source
class SheetFile:
    """One sheet on disk behind the array attributes a dataset needs (``shape``, ``dtype``); ``read_pinned()`` returns its
    pixels as a pinned tensor, decoded on a background thread started by ``prefetch()`` (or at the first call)."""

    _reader = None
    _lock = threading.Lock()

    def __init__(self, path, stack="TZ", mode="L"):
        self.path, self.stack, self.mode = str(path), stack, mode
        ext = self.path.rsplit(".", 1)[-1].lower()
        self._kind = "tiff" if ext in ("tif", "tiff") else ("czi" if ext == "czi" else "pillow")
        self._future = None
        self._eager = None
        if self._kind == "tiff":
            f, h, w, bits, _ = tiff_probe(self.path)
            self.shape, self.dtype = (f, h, w), np.dtype(np.uint8 if bits == 8 else np.uint16)
        else:                                        # geometry only known after decoding: decode now, keep the pinned copy
            self._eager = self._decode()
            self.shape = tuple(self._eager.shape)
            self.dtype = np.dtype(np.uint8 if self._eager.dtype == torch.uint8 else np.uint16)

    def _decode(self):
        if self._kind == "tiff":
            return read_tiff(self.path, pin=True)
        a = czi_to_sheet(*read_czi(self.path), stack=self.stack, mode=self.mode) if self._kind == "czi" else _pillow_stack(self.path, self.mode)
        t = torch.as_tensor(a.view(np.int16) if a.dtype == np.uint16 else a)
        return t.pin_memory() if _can_pin() else t

    @classmethod
    def _pool(cls):
        with cls._lock:
            if cls._reader is None:
                cls._reader = ThreadPoolExecutor(max_workers=2, thread_name_prefix="pssr-sheet-reader")
            return cls._reader

    def prefetch(self):
        if self._eager is None and self._future is None:
            self._future = self._pool().submit(self._decode)

    def ready(self):
        """True if ``read_pinned()`` would not block on the decode."""
        return self._eager is not None or (self._future is not None and self._future.done())

    def read_pinned(self):
        if self._eager is not None:
            return self._eager
        self.prefetch()
        t, self._future = self._future.result(), None      # the pinned buffer is handed over (not cached: sheets may exceed host RAM)
        return t


# ------------------------------------------------------------------------------------ CZI (ZISRAW)
_CZI_PIXEL = {0: (np.uint8, 1), 1: (np.uint16, 1), 2: (np.float32, 1), 3: (np.uint8, 3), 4: (np.uint16, 3)}


def read_czi(path):
    """Minimal ZISRAW reader: -> (array, axes) with one axis letter per array dimension (sub-block dimension order, e.g.
    "TZCYX0"), uncompressed sub-blocks only.  Layout: every segment starts with a 32-byte header (16-byte id, allocated size,
    used size); the file header segment points at the sub-block directory, whose entries ("DV" schema) carry the pixel type,
    the file position of the sub-block and its dimension entries (name, start, size); a sub-block segment holds its own copy of
    the entry, then metadata, then the pixel data at offset max(256, 16 + entry size) + metadata size."""
    with open(path, "rb") as f:
        data = f.read()
    if data[:10] != b"ZISRAWFILE":
        raise ValueError(f"{path} is not a CZI (ZISRAW) file")
    dir_pos = struct.unpack_from("<q", data, 32 + 36)[0]          # FileHeader: ... DirectoryPosition at payload offset 36
    if data[dir_pos:dir_pos + 15] != b"ZISRAWDIRECTORY":
        raise ValueError("CZI: sub-block directory not found")
    n_entries = struct.unpack_from("<i", data, dir_pos + 32)[0]
    p = dir_pos + 32 + 128
    entries = []
    for _ in range(n_entries):
        if data[p:p + 2] != b"DV":
            raise NotImplementedError("CZI: only DV directory entries are supported")
        pixel_type, file_pos, _fp, compression = struct.unpack_from("<iqii", data, p + 2)
        n_dims = struct.unpack_from("<i", data, p + 28)[0]
        dims = []
        for d in range(n_dims):
            q = p + 32 + 20 * d
            name = data[q:q + 4].split(b"\0")[0].decode()
            start, size = struct.unpack_from("<ii", data, q + 4)
            dims.append((name, start, size))
        entries.append((pixel_type, file_pos, compression, dims))
        p += 32 + 20 * n_dims
    if not entries:
        raise ValueError("CZI: no sub-blocks")
    if any(e[2] != 0 for e in entries):
        raise NotImplementedError("CZI: compressed sub-blocks are not supported (uncompressed only)")
    dtype, samples = _CZI_PIXEL.get(entries[0][0], (None, 0))
    if dtype is None:
        raise NotImplementedError(f"CZI: pixel type {entries[0][0]} is not supported")
    names = [d[0] for d in entries[0][3]][::-1]                   # stored fastest-varying first (X, Y, C, Z, T ...)
    lo = {n: min(d[1] for e in entries for d in e[3] if d[0] == n) for n in names}
    hi = {n: max(d[1] + d[2] for e in entries for d in e[3] if d[0] == n) for n in names}
    shape = [hi[n] - lo[n] for n in names] + ([samples] if samples > 1 else [1])
    out = np.zeros(shape, dtype=dtype)
    for pixel_type, file_pos, _c, dims in entries:
        if data[file_pos:file_pos + 14] != b"ZISRAWSUBBLOCK":
            raise ValueError("CZI: sub-block segment not found")
        q = file_pos + 32
        meta_size, _att, data_size = struct.unpack_from("<iiq", data, q)
        entry_size = 32 + 20 * len(dims)
        off = q + max(256, 16 + entry_size) + meta_size
        sub_shape = [dict((d[0], d[2]) for d in dims)[n] for n in names] + [shape[-1]]
        block = np.frombuffer(data, dtype=dtype, count=int(np.prod(sub_shape)), offset=off).reshape(sub_shape)
        sl = tuple(slice(dict((d[0], d[1]) for d in dims)[n] - lo[n], dict((d[0], d[1]) for d in dims)[n] - lo[n] + s)
                   for n, s in zip(names, sub_shape[:-1]))
        out[sl] = block
    return out, "".join(names) + "0"


def czi_to_sheet(image, axes, stack="TZ", mode="L"):
    """pssr/data.py:585-619: keep the TZCXY axes (index 0 of every other one), order them as "TZCXY", average the channels for
    mode "L", select / order the stack axes, flatten to [frames, X, Y] and scale the maximum to 255 -> uint8.
    (The reference's ``out_axes.rfind`` moves X before Y exactly as written here.)"""
    out_axes = "TZCXY"
    slice_idx, slice_axes = [], []
    for axis in axes:
        if axis not in out_axes:
            slice_idx.append(0)
        else:
            slice_idx.append(slice(None))
            slice_axes.append(axis)
    image = image[tuple(slice_idx)]
    for axis in out_axes:                       # the reference assumes every out axis exists (its TODO); missing ones get length 1
        if axis not in slice_axes:
            image = image[np.newaxis]
            slice_axes.insert(0, axis)
    axes_idx = [out_axes.rfind(axis) for axis in slice_axes]
    image = np.moveaxis(image, range(len(image.shape)), axes_idx)
    if mode == "L":
        image = np.mean(image, axis=2)
    if stack == "T":
        image = image[:, 0]
    elif stack == "Z":
        image = image[0]
    elif stack == "ZT":
        image = np.moveaxis(image, 0, 1)
    elif stack != "TZ":
        raise ValueError(f"Stack type {stack} is not valid.")
    image = np.reshape(image, [-1, image.shape[-2], image.shape[-1]])
    if image.max() != 0:
        image = image / (image.max() / 255)
    return image.astype(np.uint8)
