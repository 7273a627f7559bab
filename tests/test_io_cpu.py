"""CPU checks of the file I/O edges (pssr2_b200/io.py + csrc/tiff_io.cu): TIFF decode / encode against Pillow
in both directions, hand-built big-endian / BigTIFF / ImageJ-hyperstack files, and the CZI path against the reference's own
`_load_sheet` (pssr/data.py:585-619) driven through a stand-in for czifile."""
import struct

import numpy as np
import pytest

from pssr2_b200 import io


def _stack(dt, shape=(3, 37, 53), seed=0):
    return np.random.default_rng(seed).integers(0, 60000 if dt == np.uint16 else 255, shape).astype(dt)


@pytest.mark.parametrize("dt", [np.uint8, np.uint16])
def test_tiff_roundtrip_and_pillow_interop(tmp_path, dt):
    from PIL import Image
    a = _stack(dt)
    io.write_tiff(tmp_path / "a.tif", a)
    assert io.tiff_probe(tmp_path / "a.tif") == (3, 37, 53, 8 * a.itemsize, True)
    assert np.array_equal(io.read_tiff(tmp_path / "a.tif"), a)
    t = io.read_tiff(tmp_path / "a.tif", pin=False)
    assert t.dtype == dt
    im = Image.open(tmp_path / "a.tif")                    # Pillow reads what we wrote
    frames = []
    for i in range(im.n_frames):
        im.seek(i)
        frames.append(np.asarray(im))
    assert np.array_equal(np.stack(frames), a)
    ims = [Image.fromarray(x) for x in a]                  # we read what Pillow wrote: strips, and the LZW fallback
    ims[0].save(tmp_path / "p.tif", format="TIFF", save_all=True, append_images=ims[1:])
    assert np.array_equal(io.read_tiff(tmp_path / "p.tif"), a)
    ims[0].save(tmp_path / "z.tif", format="TIFF", save_all=True, append_images=ims[1:], compression="tiff_lzw")
    assert io.tiff_probe(tmp_path / "z.tif")[4] is False
    assert np.array_equal(io.read_tiff(tmp_path / "z.tif"), a)
    io.write_tiff(tmp_path / "single.tif", a[0])           # 2-D input -> one frame
    assert io.read_tiff(tmp_path / "single.tif").shape == (1, 37, 53)
    with pytest.raises(TypeError):
        io.write_tiff(tmp_path / "f.tif", a.astype(np.float32))
    with pytest.raises(RuntimeError):
        io.read_tiff(tmp_path / "missing.tif")


def _handmade_tiff(path, a, big_endian=False, bigtiff=False, imagej=False, strips=1):
    """Classic / BigTIFF writer independent of the library's: multi-strip frames, either byte order, ImageJ contiguous stacks."""
    e = ">" if big_endian else "<"
    frames, h, w = a.shape
    bits = a.itemsize * 8
    data = a.astype(a.dtype.newbyteorder(e)).tobytes()
    fb = h * w * a.itemsize
    out = bytearray((b"MM" if big_endian else b"II") + struct.pack(e + "H", 43 if bigtiff else 42))
    if bigtiff:
        out += struct.pack(e + "HHQ", 8, 0, 0)
    else:
        out += struct.pack(e + "I", 0)
    data_off = len(out)
    out += data
    if len(out) % 2:
        out += b"\0"
    desc = b"ImageJ=1.53\nimages=%d\nslices=%d\n\0" % (frames, frames)
    desc_off = len(out)
    if imagej:
        out += desc + (b"\0" if len(desc) % 2 else b"")
    rows = -(-h // strips)
    n_ifd = 1 if imagej else frames
    ifd_offs = []
    for i in range(n_ifd):
        offs = [data_off + i * fb + s * rows * w * a.itemsize for s in range(strips)]
        cnts = [min(rows, h - s * rows) * w * a.itemsize for s in range(strips)]
        if imagej:
            offs, cnts = [data_off], [fb]
        fmt_l, ty_l, sz_l = ("Q", 16, 8) if bigtiff else ("I", 4, 4)
        arr_off = len(out)
        if len(offs) > 1:
            out += struct.pack(e + fmt_l * len(offs), *offs) + struct.pack(e + fmt_l * len(cnts), *cnts)
        ents = [(256, 3, 1, w), (257, 3, 1, h), (258, 3, 1, bits), (259, 3, 1, 1), (262, 3, 1, 1),
                (273, ty_l, len(offs), offs[0] if len(offs) == 1 else arr_off), (277, 3, 1, 1), (278, 3, 1, rows if not imagej else h),
                (279, ty_l, len(cnts), cnts[0] if len(cnts) == 1 else arr_off + sz_l * len(offs))]
        if imagej:
            ents.insert(5, (270, 2, len(desc), desc_off))
        ifd_offs.append(len(out))
        out += struct.pack(e + ("Q" if bigtiff else "H"), len(ents))
        for tag, ty, cnt, val in ents:
            out += struct.pack(e + "HH", tag, ty) + struct.pack(e + ("Q" if bigtiff else "I"), cnt)
            vb = 8 if bigtiff else 4
            if ty == 3 and cnt == 1:
                out += struct.pack(e + "H", val) + b"\0" * (vb - 2)
            else:
                out += struct.pack(e + ("Q" if bigtiff else "I"), val)
        out += struct.pack(e + ("Q" if bigtiff else "I"), 0)           # next IFD, patched below
    ptr = 8 if bigtiff else 4
    struct.pack_into(e + ("Q" if bigtiff else "I"), out, 8 if bigtiff else 4, ifd_offs[0])
    for i in range(n_ifd - 1):
        nxt_pos = ifd_offs[i + 1] - (0)      # the "next" field is the last word of IFD i: right before the arrays / IFD that follow
        n_ents = 9
        end_of_ifd = ifd_offs[i] + (8 if bigtiff else 2) + n_ents * (20 if bigtiff else 12)
        struct.pack_into(e + ("Q" if bigtiff else "I"), out, end_of_ifd, ifd_offs[i + 1])
    with open(path, "wb") as f:
        f.write(out)


@pytest.mark.parametrize("kw", [dict(big_endian=True), dict(bigtiff=True), dict(strips=4), dict(big_endian=True, bigtiff=True, strips=3),
                                dict(imagej=True), dict(imagej=True, big_endian=True)])
@pytest.mark.parametrize("dt", [np.uint8, np.uint16])
def test_tiff_layout_variants(tmp_path, kw, dt):
    a = _stack(dt, (4, 21, 30), seed=3)
    _handmade_tiff(tmp_path / "v.tif", a, **kw)
    f, h, w, bits, native = io.tiff_probe(tmp_path / "v.tif")
    assert (f, h, w, bits, native) == (4, 21, 30, 8 * a.itemsize, True)
    assert np.array_equal(io.read_tiff(tmp_path / "v.tif"), a)


def test_sheet_file_is_lazy_and_pinned(tmp_path):
    a = _stack(np.uint16, (2, 64, 48), seed=5)
    io.write_tiff(tmp_path / "s.tif", a)
    s = io.SheetFile(tmp_path / "s.tif")
    assert s.shape == (2, 64, 48) and s.dtype == np.uint16 and not s.ready()
    s.prefetch()
    t = s.read_pinned()
    import torch
    assert t.is_pinned() == torch.cuda.is_available() and np.array_equal(t.numpy().view(np.uint16), a)
    with io.TiffWriter(threads=2) as wr:
        for i in range(5):
            wr.write(tmp_path / f"w{i}.tif", a[i % 2])
    assert all(np.array_equal(io.read_tiff(tmp_path / f"w{i}.tif")[0], a[i % 2]) for i in range(5))


# ------------------------------------------------------------------------------------- CZI
def _write_czi(path, arr, axes):
    """Test-only ZISRAW writer: one uncompressed sub-block per (T, Z, C) plane; `arr` is indexed by `axes` (e.g. "TZCYX")."""
    assert axes[-2:] == "YX"
    pix = {np.dtype(np.uint8): 0, np.dtype(np.uint16): 1}[arr.dtype]

    def seg(sid, payload):
        return sid.ljust(16, b"\0") + struct.pack("<qq", len(payload), len(payload)) + payload

    lead = axes[:-2]
    planes = list(np.ndindex(*arr.shape[:-2]))
    H, W = arr.shape[-2:]

    def entry(idx, pos):
        dims = [("X", 0, W), ("Y", 0, H)] + [(ax, idx[k], 1) for k, ax in reversed(list(enumerate(lead)))]
        b = b"DV" + struct.pack("<iqii", pix, pos, 0, 0) + b"\0" * 6 + struct.pack("<i", len(dims))
        for nme, st, sz in dims:
            b += nme.encode().ljust(4, b"\0") + struct.pack("<ii", st, sz) + struct.pack("<f", 0.0) + struct.pack("<i", sz)
        return b

    header_payload_size = 512
    pos = 32 + header_payload_size
    blocks, entries = [], []
    for idx in planes:
        e = entry(idx, pos)
        head = struct.pack("<iiq", 0, 0, H * W * arr.itemsize) + e
        head = head.ljust(max(256, 16 + len(e)), b"\0")
        payload = head + arr[idx].tobytes()
        s = seg(b"ZISRAWSUBBLOCK", payload)
        entries.append(e)
        blocks.append(s)
        pos += len(s)
    dir_pos = pos
    directory = seg(b"ZISRAWDIRECTORY", struct.pack("<i", len(entries)).ljust(128, b"\0") + b"".join(entries))
    hp = bytearray(header_payload_size)
    struct.pack_into("<ii", hp, 0, 1, 0)
    struct.pack_into("<q", hp, 36, dir_pos)
    with open(path, "wb") as f:
        f.write(seg(b"ZISRAWFILE", bytes(hp)) + b"".join(blocks) + directory)


@pytest.mark.parametrize("stack", ["TZ", "ZT", "T", "Z"])
@pytest.mark.parametrize("dt", [np.uint8, np.uint16])
def test_czi_decode_and_sheet_rule(tmp_path, stack, dt):
    rng = np.random.default_rng(7)
    arr = rng.integers(0, 4000 if dt == np.uint16 else 200, (2, 3, 2, 12, 17)).astype(dt)     # T Z C Y X
    _write_czi(tmp_path / "a.czi", arr, "TZCYX")
    got, axes = io.read_czi(tmp_path / "a.czi")
    assert axes == "TZCYX0" and got.shape == (2, 3, 2, 12, 17, 1) and np.array_equal(got[..., 0], arr)
    sheet = io.czi_to_sheet(got, axes, stack=stack, mode="L")
    # independent statement of pssr/data.py:585-619 on the (T, Z, C, X, Y)-ordered array
    im = np.moveaxis(arr, (3, 4), (4, 3)).astype(np.float64).mean(axis=2)           # "TZCXY": X before Y, channel mean
    im = {"TZ": im, "ZT": np.moveaxis(im, 0, 1), "T": im[:, 0], "Z": im[0]}[stack]
    im = im.reshape(-1, im.shape[-2], im.shape[-1])
    want = (im / (im.max() / 255)).astype(np.uint8)
    assert sheet.dtype == np.uint8 and np.array_equal(sheet, want)
    s = io.SheetFile(tmp_path / "a.czi", stack=stack)
    assert s.shape == want.shape and np.array_equal(s.read_pinned().numpy(), want)


def test_czi_sheet_rule_matches_reference(tmp_path):
    """`czi_to_sheet` against the reference's own `_load_sheet` (pssr/data.py:576-619), with czifile replaced by a stand-in that
    serves the array `read_czi` decoded."""
    from oracle.refshim import import_reference, reference_available
    if not reference_available():
        pytest.skip("reference not present")
    import sys
    import_reference()
    from pssr import data as RD
    arr = np.random.default_rng(9).integers(0, 3000, (2, 2, 3, 10, 14)).astype(np.uint16)
    _write_czi(tmp_path / "r.czi", arr, "TZCYX")
    got, axes = io.read_czi(tmp_path / "r.czi")

    class FakeCzi:
        def __init__(self, p):
            self.axes = axes
        def asarray(self):
            return got
    old = RD.czifile.CziFile
    RD.czifile.CziFile = FakeCzi
    try:
        for stack in ("TZ", "ZT", "T", "Z"):
            want = RD._load_sheet(str(tmp_path), "r.czi", stack, "L")
            assert np.array_equal(io.czi_to_sheet(got, axes, stack=stack, mode="L"), want), stack
    finally:
        RD.czifile.CziFile = old
