// File I/O edges of the hot path: baseline TIFF decode straight into caller-owned (pinned) host memory and
// TIFF encode of predictions / stitched sheets.  Replaces the reference's tifffile.imread / tifffile.imwrite call sites
// (pssr/data.py:566-571, :621-625, pssr/predict.py:71, pssr/util.py:103) for the layouts microscopy sheets come in:
// grayscale 8 / 16 bit, uncompressed strips, one IFD per frame (classic TIFF or BigTIFF, either byte order) or an ImageJ
// hyperstack whose frames follow the first one contiguously.  Anything else (compression, tiles, RGB) reports `native = 0` and
// the host mirror (pssr2_b200/io.py) decodes it with Pillow instead.  Host code only: no kernel in this file.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include "common.cuh"

namespace pssr {

struct TiffFrame {
  int64_t w = 0, h = 0, bits = 8, spp = 1, compression = 1, planar = 1, rows_per_strip = 0, photometric = 1;
  bool tiled = false;
  std::vector<uint64_t> offsets, counts;
};

struct TiffFile {
  FILE* f = nullptr;
  bool big_endian = false, bigtiff = false;
  std::vector<TiffFrame> frames;
  int64_t imagej_images = 0;
  ~TiffFile() { if (f) fclose(f); }
};

static uint64_t rd(const uint8_t* p, int n, bool be) {
  uint64_t v = 0;
  for (int i = 0; i < n; ++i) v |= (uint64_t)p[i] << (8 * (be ? n - 1 - i : i));
  return v;
}

static int type_size(int t) {
  switch (t) {
    case 1: case 2: case 6: case 7: return 1;
    case 3: case 8: return 2;
    case 4: case 9: case 11: case 13: return 4;
    case 5: case 10: case 12: case 16: case 17: case 18: return 8;
    default: return 0;
  }
}

// reads `count` integers of TIFF type `type` from the entry's value field (inline) or from the file offset it holds
static bool read_values(TiffFile& t, int type, uint64_t count, const uint8_t* valfield, int valbytes, std::vector<uint64_t>& out) {
  const int ts = type_size(type);
  if (ts == 0 || count > (1ull << 28)) return false;
  std::vector<uint8_t> buf((size_t)(ts * count));
  if (ts * count <= (uint64_t)valbytes) {
    memcpy(buf.data(), valfield, buf.size());
  } else {
    const uint64_t off = rd(valfield, valbytes, t.big_endian);
    if (fseeko(t.f, (off_t)off, SEEK_SET) != 0 || fread(buf.data(), 1, buf.size(), t.f) != buf.size()) return false;
  }
  out.resize((size_t)count);
  for (uint64_t i = 0; i < count; ++i) out[(size_t)i] = rd(buf.data() + i * ts, ts, t.big_endian);
  return true;
}

static int tiff_open(const char* path, TiffFile& t) {
  t.f = fopen(path, "rb");
  PSSR_REQUIRE(t.f != nullptr, PSSR_EINVAL, "tiff: cannot open %s", path);
  uint8_t hdr[16];
  PSSR_REQUIRE(fread(hdr, 1, 8, t.f) == 8, PSSR_EINVAL, "tiff: %s is too short", path);
  PSSR_REQUIRE((hdr[0] == 'I' && hdr[1] == 'I') || (hdr[0] == 'M' && hdr[1] == 'M'), PSSR_EINVAL, "tiff: %s has no TIFF byte-order mark", path);
  t.big_endian = hdr[0] == 'M';
  const uint64_t magic = rd(hdr + 2, 2, t.big_endian);
  PSSR_REQUIRE(magic == 42 || magic == 43, PSSR_EINVAL, "tiff: %s has magic %d", path, (int)magic);
  t.bigtiff = magic == 43;
  uint64_t ifd;
  if (t.bigtiff) {
    PSSR_REQUIRE(fread(hdr + 8, 1, 8, t.f) == 8, PSSR_EINVAL, "tiff: truncated BigTIFF header");
    ifd = rd(hdr + 8, 8, t.big_endian);
  } else {
    ifd = rd(hdr + 4, 4, t.big_endian);
  }
  const int esz = t.bigtiff ? 20 : 12, cntb = t.bigtiff ? 8 : 2, valb = t.bigtiff ? 8 : 4;
  int guard = 0;
  while (ifd != 0 && guard++ < (1 << 20)) {
    uint8_t nb[8];
    PSSR_REQUIRE(fseeko(t.f, (off_t)ifd, SEEK_SET) == 0 && fread(nb, 1, cntb, t.f) == (size_t)cntb, PSSR_EINVAL, "tiff: bad IFD offset");
    const uint64_t n = rd(nb, cntb, t.big_endian);
    PSSR_REQUIRE(n > 0 && n < 4096, PSSR_EINVAL, "tiff: implausible IFD entry count");
    std::vector<uint8_t> ent((size_t)(n * esz + valb));
    PSSR_REQUIRE(fread(ent.data(), 1, ent.size(), t.f) == ent.size(), PSSR_EINVAL, "tiff: truncated IFD");
    TiffFrame fr;
    for (uint64_t i = 0; i < n; ++i) {
      const uint8_t* e = ent.data() + i * esz;
      const int tag = (int)rd(e, 2, t.big_endian), type = (int)rd(e + 2, 2, t.big_endian);
      const uint64_t count = rd(e + 4, t.bigtiff ? 8 : 4, t.big_endian);
      const uint8_t* val = e + (t.bigtiff ? 12 : 8);
      std::vector<uint64_t> v;
      const bool wanted = tag == 256 || tag == 257 || tag == 258 || tag == 259 || tag == 262 || tag == 273 || tag == 277 || tag == 278 ||
                          tag == 279 || tag == 284 || tag == 322 || tag == 324;
      if (tag == 270 && t.frames.empty() && type == 2 && count < (1u << 20)) {       // ImageDescription: ImageJ hyperstack marker
        std::vector<uint8_t> s((size_t)count + 1, 0);
        if (count <= (uint64_t)valb) memcpy(s.data(), val, (size_t)count);
        else {
          const uint64_t off = rd(val, valb, t.big_endian);
          const off_t here = ftello(t.f);
          if (fseeko(t.f, (off_t)off, SEEK_SET) == 0) { size_t got = fread(s.data(), 1, (size_t)count, t.f); (void)got; }
          fseeko(t.f, here, SEEK_SET);
        }
        const char* d = reinterpret_cast<const char*>(s.data());
        const char* im = strstr(d, "images=");
        if (strstr(d, "ImageJ=") != nullptr && im != nullptr) t.imagej_images = atoll(im + 7);
        continue;
      }
      if (!wanted) continue;
      const off_t here = ftello(t.f);
      const bool ok = read_values(t, type, count, val, valb, v);
      fseeko(t.f, here, SEEK_SET);
      PSSR_REQUIRE(ok && !v.empty(), PSSR_EINVAL, "tiff: unreadable tag %d", tag);
      switch (tag) {
        case 256: fr.w = (int64_t)v[0]; break;
        case 257: fr.h = (int64_t)v[0]; break;
        case 258: fr.bits = (int64_t)v[0]; break;
        case 259: fr.compression = (int64_t)v[0]; break;
        case 262: fr.photometric = (int64_t)v[0]; break;
        case 273: fr.offsets = v; break;
        case 277: fr.spp = (int64_t)v[0]; break;
        case 278: fr.rows_per_strip = (int64_t)v[0]; break;
        case 279: fr.counts = v; break;
        case 284: fr.planar = (int64_t)v[0]; break;
        case 322: case 324: fr.tiled = true; break;
      }
    }
    t.frames.push_back(fr);
    ifd = rd(ent.data() + n * esz, valb, t.big_endian);
  }
  PSSR_REQUIRE(!t.frames.empty(), PSSR_EINVAL, "tiff: %s holds no image", path);
  return PSSR_OK;
}

static bool frame_native(const TiffFrame& f) {
  return !f.tiled && f.compression == 1 && f.spp == 1 && (f.bits == 8 || f.bits == 16) && f.w > 0 && f.h > 0 && !f.offsets.empty() &&
         f.offsets.size() == f.counts.size() && f.photometric <= 1;
}

}  // namespace pssr

using namespace pssr;

extern "C" {

int pssr_tiff_probe(const char* path, int32_t* frames, int32_t* h, int32_t* w, int32_t* bits, int32_t* native) {
  PSSR_REQUIRE(path && frames && h && w && bits && native, PSSR_EINVAL, "tiff_probe: null argument");
  TiffFile t;
  int rc = tiff_open(path, t);
  if (rc != PSSR_OK) return rc;
  const TiffFrame& f0 = t.frames[0];
  bool nat = true;
  for (const TiffFrame& f : t.frames) nat = nat && frame_native(f) && f.w == f0.w && f.h == f0.h && f.bits == f0.bits;
  int64_t n = (int64_t)t.frames.size();
  if (nat && n == 1 && t.imagej_images > 1 && f0.offsets.size() == 1) n = t.imagej_images;     // contiguous ImageJ hyperstack
  *frames = (int32_t)n; *h = (int32_t)f0.h; *w = (int32_t)f0.w; *bits = (int32_t)f0.bits; *native = nat ? 1 : 0;
  return PSSR_OK;
}

int pssr_tiff_read(const char* path, void* dst, int64_t dst_bytes) {
  PSSR_REQUIRE(path && dst, PSSR_EINVAL, "tiff_read: null argument");
  TiffFile t;
  int rc = tiff_open(path, t);
  if (rc != PSSR_OK) return rc;
  const TiffFrame& f0 = t.frames[0];
  const int64_t bpp = f0.bits / 8, frame_bytes = f0.w * f0.h * bpp;
  for (const TiffFrame& f : t.frames)
    PSSR_REQUIRE(frame_native(f) && f.w == f0.w && f.h == f0.h && f.bits == f0.bits, PSSR_EUNSUP, "tiff_read: %s is not a baseline grayscale stack", path);
  uint8_t* out = reinterpret_cast<uint8_t*>(dst);
  int64_t n = (int64_t)t.frames.size();
  if (n == 1 && t.imagej_images > 1 && f0.offsets.size() == 1) {
    n = t.imagej_images;
    PSSR_REQUIRE(dst_bytes >= n * frame_bytes, PSSR_EINVAL, "tiff_read: destination too small");
    PSSR_REQUIRE(fseeko(t.f, (off_t)f0.offsets[0], SEEK_SET) == 0 && fread(out, 1, (size_t)(n * frame_bytes), t.f) == (size_t)(n * frame_bytes),
                 PSSR_EINVAL, "tiff_read: truncated ImageJ stack");
  } else {
    PSSR_REQUIRE(dst_bytes >= n * frame_bytes, PSSR_EINVAL, "tiff_read: destination too small");
    for (int64_t i = 0; i < n; ++i) {
      const TiffFrame& f = t.frames[(size_t)i];
      int64_t done = 0;
      for (size_t s = 0; s < f.offsets.size() && done < frame_bytes; ++s) {
        int64_t want = (int64_t)f.counts[s];
        if (want > frame_bytes - done) want = frame_bytes - done;
        PSSR_REQUIRE(fseeko(t.f, (off_t)f.offsets[s], SEEK_SET) == 0 && fread(out + i * frame_bytes + done, 1, (size_t)want, t.f) == (size_t)want,
                     PSSR_EINVAL, "tiff_read: truncated strip");
        done += want;
      }
      PSSR_REQUIRE(done == frame_bytes, PSSR_EINVAL, "tiff_read: frame %d has %lld of %lld bytes", (int)i, (long long)done, (long long)frame_bytes);
    }
  }
  if (bpp == 2 && t.big_endian) {        // to native (little-endian) order
    uint16_t* p = reinterpret_cast<uint16_t*>(out);
    const int64_t cnt = n * f0.w * f0.h;
    for (int64_t i = 0; i < cnt; ++i) p[i] = (uint16_t)((p[i] >> 8) | (p[i] << 8));
  }
  if (f0.photometric == 0) {             // WhiteIsZero
    const int64_t cnt = n * f0.w * f0.h;
    if (bpp == 1) for (int64_t i = 0; i < cnt; ++i) out[i] = (uint8_t)(255 - out[i]);
    else { uint16_t* p = reinterpret_cast<uint16_t*>(out); for (int64_t i = 0; i < cnt; ++i) p[i] = (uint16_t)(65535 - p[i]); }
  }
  return PSSR_OK;
}

// [frames][h][w] uint8 / uint16 -> little-endian TIFF, one uncompressed strip and one IFD per frame; BigTIFF beyond 4 GB.
int pssr_tiff_write(const char* path, const void* src, int32_t frames, int32_t h, int32_t w, int32_t bits) {
  PSSR_REQUIRE(path && src && frames >= 1 && h >= 1 && w >= 1 && (bits == 8 || bits == 16), PSSR_EINVAL, "tiff_write: bad arguments");
  const uint64_t data_bytes = (uint64_t)h * w * (bits / 8);
  const uint64_t frame_bytes = data_bytes + (data_bytes & 1);      // IFDs start on even offsets
  const bool big = frame_bytes * frames + (uint64_t)frames * 256 + 16 >= 0xFFFF0000ull;
  FILE* f = fopen(path, "wb");
  PSSR_REQUIRE(f != nullptr, PSSR_EINVAL, "tiff_write: cannot create %s", path);
  auto put = [&](uint64_t v, int n) { uint8_t b[8]; for (int i = 0; i < n; ++i) b[i] = (uint8_t)(v >> (8 * i)); return fwrite(b, 1, n, f) == (size_t)n; };
  const int nent = 9, esz = big ? 20 : 12, offb = big ? 8 : 4;
  const uint64_t hdr = big ? 16 : 8;
  const uint64_t ifd_bytes = (big ? 8 : 2) + (uint64_t)nent * esz + offb;
  bool ok = fwrite("II", 1, 2, f) == 2 && put(big ? 43 : 42, 2);
  if (big) ok = ok && put(8, 2) && put(0, 2) && put(hdr + frame_bytes, 8);
  else ok = ok && put(hdr + frame_bytes, 4);
  uint64_t pos = hdr;
  const uint8_t* p = reinterpret_cast<const uint8_t*>(src);
  for (int i = 0; i < frames && ok; ++i) {
    ok = fwrite(p + (uint64_t)i * data_bytes, 1, data_bytes, f) == data_bytes && (frame_bytes == data_bytes || put(0, 1));
    const uint64_t data_off = pos;
    pos += frame_bytes + ifd_bytes;
    const uint64_t next = i + 1 < frames ? pos + frame_bytes : 0;
    auto entry = [&](int tag, int type, uint64_t value) {
      bool r = put((uint64_t)tag, 2) && put((uint64_t)type, 2) && put(1, big ? 8 : 4);
      return r && put(value, offb);
    };
    ok = ok && put(nent, big ? 8 : 2);
    const int lt = big ? 16 : 4;       // LONG8 in BigTIFF, LONG otherwise
    ok = ok && entry(256, 4, (uint64_t)w) && entry(257, 4, (uint64_t)h) && entry(258, 3, (uint64_t)bits) && entry(259, 3, 1) && entry(262, 3, 1) &&
         entry(273, lt, data_off) && entry(277, 3, 1) && entry(278, 4, (uint64_t)h) && entry(279, lt, data_bytes);
    ok = ok && put(next, offb);
  }
  ok = (fclose(f) == 0) && ok;
  PSSR_REQUIRE(ok, PSSR_EINVAL, "tiff_write: short write to %s", path);
  return PSSR_OK;
}

}  // extern "C"
