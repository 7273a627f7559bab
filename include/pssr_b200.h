/*
 * pssr_b200.h -- C ABI of libpssr_b200.so: the B200 (sm_100a) kernels behind PSSR2's
 * test/predict hot path (crappify -> SR network forward -> stitch -> score).
 *
 * The reference (ucsdmanorlab/PSSR2 v2.4.0) is pure Python and has no FFI; each entry
 * point below names the reference call site (file:line under /root/reference) whose
 * arithmetic it replaces.  INTEGRATION.md shows the ctypes binding a PSSR2 maintainer
 * would add.
 *
 * Conventions
 *  - every function returns 0 on success, a negative PSSR_E* code on failure; the message
 *    is available (thread-local) from pssr_last_error().
 *  - all pointers are DEVICE pointers owned by the caller unless the name says host_;
 *    the library never frees caller memory and allocates only inside an opaque plan
 *    (pssr_plan_create / pssr_plan_destroy) and, once per device, a few constant tables
 *    (resample coefficients per geometry, Poisson alias tables: < 1 MB).
 *  - `stream` is a cudaStream_t passed as void*; calls are asynchronous w.r.t. the host.
 *  - images are row-major, innermost dimension last; shapes are explicit.
 */
#ifndef PSSR_B200_H
#define PSSR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PSSR_OK 0
#define PSSR_EINVAL (-1)   /* bad argument                        */
#define PSSR_ECUDA (-2)    /* CUDA runtime / driver error         */
#define PSSR_EUNSUP (-3)   /* shape or mode outside the supported set */

const char* pssr_last_error(void);
/* Library version string, e.g. "pssr_b200 0.1 sm_100a". */
const char* pssr_version(void);
/* Number of kernels this library has launched so far in this process (bench.py's
 * gpu_launches counter). */
int64_t pssr_launch_count(void);
/* Developer diagnostic (no reference counterpart): copies the convolution kernels' per-CTA clock64 timeline
 * (recorded only while the environment variable PSSR_DBG has bit 16 set) into `out` (n <= 148*128 int64). */
int pssr_debug_trace(int64_t* out, int64_t n);

/* ------------------------------------------------------------------------------------
 * Family 1: fused crappify  (pssr/data.py:471-495 `_gen_pair`, :629-638 `_sliding_window`,
 * :536-551 `_square_crop`/`_pad_image`, Pillow `Image.resize(BILINEAR)` at :483,
 * pssr/crappifiers.py:26-105, round/clip at data.py:487).
 *
 * One launch turns `n_tiles` HR tiles, addressed inside resident sheets, into LR tiles:
 *   tile gather (+ reflect pad) -> Pillow-exact two-pass antialiased triangle downscale
 *   (uint8: 22-bit fixed point, uint16: double) -> noise stages -> round-half-even ->
 *   clip [0,255] -> float32 (and optionally the network's u8 staging copy).
 * ------------------------------------------------------------------------------------ */

/* noise stage kinds, applied in order (MultiCrappifier, crappifiers.py:38-43) */
#define PSSR_NOISE_POISSON 1   /* crappifiers.py:82-86 */
#define PSSR_NOISE_GAUSSIAN 2  /* crappifiers.py:62-64 */
#define PSSR_NOISE_SALTPEPPER 3/* crappifiers.py:103-105 */

/* noise sources */
#define PSSR_RNG_INJECTED 0    /* draws supplied by the caller (bit-exact parity mode) */
#define PSSR_RNG_PHILOX 1      /* on-device counter-based Philox4x32-10, keyed by (seed, tile, pixel) */

typedef struct {
  int32_t kind;          /* PSSR_NOISE_*                                                  */
  int32_t rng;           /* PSSR_RNG_*                                                    */
  double intensity;      /* already resolved for this call (spread draw done by the host):
                            POISSON mix i; GAUSSIAN sigma and SALTPEPPER amount (Philox only) */
  double gain;
  int32_t mix_in_f32;    /* POISSON: 1 = x*(1-i) is evaluated in float32 (python-scalar intensity),
                            0 = in float64 (np.float64 intensity, i.e. spread > 0) -- NumPy>=2 promotion */
  int32_t reserved;
  /* injected draws, one value per LR pixel of every tile, layout [n_tiles][frames][lr][lr]:
   *   POISSON:    int64 samples  y  (np.random.poisson output)
   *   GAUSSIAN:   double samples g  (np.random.normal(gain, intensity) output, i.e. already
   *               shifted/scaled)
   *   SALTPEPPER: uint8 bit0 = flipped, bit1 = salted                                     */
  const void* injected;
} pssr_noise_stage_t;

typedef struct {
  /* source sheets: `sheets` points at a DEVICE array of n_sheets device pointers; every sheet is
   * [frames_total][sheet_h][sheet_w], dtype uint8 (elem_bytes=1) or uint16 (elem_bytes=2).
   * Allocation contract: 16-byte-aligned sheets are staged with aligned 16-byte vector copies, which may read up to 15 bytes
   * past the last pixel of a sheet (never past its 16-byte granule): the allocation must extend to the next multiple of 16
   * bytes, as cudaMalloc / torch allocations always do.  Misaligned sheets take the element-wise path. */
  const void* const* sheets;
  int32_t n_sheets;
  int32_t elem_bytes;
  int32_t sheet_h, sheet_w;
  /* per tile (device int32 arrays of length n_tiles): sheet index, first frame, top row,
   * left column of the valid region, and its valid height/width (< hr_res => reflect pad on the
   * bottom/right as np.pad(..., "reflect") does, data.py:548-551) */
  const int32_t* tile_sheet;
  const int32_t* tile_frame;
  const int32_t* tile_y;
  const int32_t* tile_x;
  const int32_t* tile_vh;
  const int32_t* tile_vw;
  int32_t n_tiles;
  int32_t frames;        /* frames in the tile's window = max(n_frames); ALL are crappified */
  int32_t hr_res;        /* HR tile edge                                        */
  int32_t lr_scale;      /* integer downscale factor                            */
  /* noise: n_stages == 0 means crappifier=None (no round/clip either, data.py:484)      */
  pssr_noise_stage_t stages[4];
  int32_t n_stages;
  int32_t clip_between;  /* MultiCrappifier(clip=True), crappifiers.py:41-42            */
  uint64_t seed;         /* Philox key                                                  */
  uint64_t tile_index0;  /* global index of tile 0 (so results do not depend on #GPUs)  */
  /* outputs */
  float* lr_out;         /* [n_tiles][lr_frames][lr][lr] float32 (what _tensor_ready yields) */
  float* hr_out;         /* optional [n_tiles][hr_frames][hr][hr] float32 HR tiles (raw values,
                            data.py:495), or NULL                                         */
  uint8_t* hr_u8_out;    /* optional [n_tiles][hr][hr] uint8 = trunc(clip(centre HR frame,0,255))
                            (`_pred_array(hr)`, predict.py:185,245-246), or NULL          */
  int32_t hr_frame0;     /* first frame (relative to the tile's frame window) and count of */
  int32_t hr_frames;     /* the HR frames kept by _slice_center (data.py:489-491)          */
  int32_t lr_frame0;     /* first LR frame kept by _slice_center (data.py:492-493) ...     */
  int32_t lr_frames;     /* ... and how many (= frames when n_frames[0] == n_frames[1])    */
  /* heterogeneous sheets: optional DEVICE int32 arrays [n_sheets] with each sheet's height / width; when NULL every sheet
   * is sheet_h x sheet_w (the reference datasets accept images of different sizes, data.py:536-551) */
  const int32_t* sheet_hs;
  const int32_t* sheet_ws;
  /* Training-time augmentation (pssr/data.py:476-480, drawn per item at :108 / :244): optional per-tile transform of the
   * cropped + reflect-padded hr_res x hr_res tile BEFORE the downscale, bit 0 = np.rot90(axes=(1,2)), then bit 1 = flip of axis 1
   * (rows), bit 2 = flip of axis 2 (columns).  The HR outputs carry the same transform.  NULL = none (the predict path).        */
  const int32_t* tile_xf;
} pssr_crappify_args_t;

int pssr_crappify(const pssr_crappify_args_t* args, void* stream);
/* The tile table of a batch (a few KB: sheet pointers + the int32 columns above) reaches the device through a one-CTA kernel
 * that reads the PINNED host buffer directly (unified addressing) instead of a cudaMemcpyAsync: a copy-engine transfer queues
 * behind every bulk sheet upload already handed to the host->device engine (measured: the second batch of a 50-stack dataset
 * waited 22 ms), a kernel on the compute stream does not.  `pinned_src` must be page-locked (cudaHostAlloc / cudaHostRegister)
 * and stay valid until the stream has passed this call; `bytes` a multiple of 4.  The ingest side of
 * pssr/data.py:471-495 (`_gen_pair` indexes host arrays; here the indices travel to the device).                              */
int pssr_table_fetch(void* dst, const void* pinned_src, int64_t bytes, void* stream);
/* The operator interface `Crappifier.crappify(image)` (pssr/crappifiers.py:13-24): the noise chain
 * alone on an arbitrary float32 / float64 array of n elements -- no downscale, no final round/clip.
 * Output is float64 (exactly representable when the reference would return float32). */
int pssr_noise_chain(const void* in, int32_t in_is_f64, double* out, int64_t n,
                     const pssr_noise_stage_t* stages, int32_t n_stages, int32_t clip_between,
                     uint64_t seed, void* stream);
/* The resample stage alone (Pillow parity tests): src [n][h][w] -> dst [n][h/scale][w/scale],
 * same dtype (uint8 / uint16) as Pillow returns before `.astype(np.float32)`. */
int pssr_resize_bilinear(const void* src, void* dst, int32_t n, int32_t h, int32_t w,
                         int32_t scale, int32_t elem_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * Family 2: network forward as a plan of fused ops  (pssr/models/resunet.py:65-96,
 * _blocks.py:15-18,39-41, rdresunet.py:104-130, _rdnet.py:95-206).
 * The host mirror (pssr2_b200/models.py) folds BatchNorm into the weights, packs them
 * K-major, and describes the forward as a list of ops; the plan owns only TMA descriptors.
 * Activations are NHWC 16-bit (bf16 or fp16, one format per plan) in caller-owned HBM.
 * ------------------------------------------------------------------------------------ */
#define PSSR_DT_BF16 0
#define PSSR_DT_FP16 1

#define PSSR_ACT_NONE 0
#define PSSR_ACT_RELU 1
#define PSSR_ACT_GELU 2      /* exact erf GELU, _rdnet.py:186 */

#define PSSR_OP_CONV 1        /* implicit-GEMM convolution on tcgen05 tensor cores          */
#define PSSR_OP_PREP 2        /* x/128-1 -> BatchNorm(eval) -> 3x3 im2col of the input      */
#define PSSR_OP_MAXPOOL 3     /* 2x2 max pool, NHWC                                         */
#define PSSR_OP_TAIL 4        /* Reconstruction.conv + *128+128 (+ clip/trunc to u8)        */
#define PSSR_OP_DWCONV_LN 5   /* depthwise 7x7 + LayerNorm2d over C (RDNet Block)           */
#define PSSR_OP_LAYERNORM 6   /* LayerNorm2d over C                                         */
#define PSSR_OP_ESE 7         /* EffectiveSE gate + layer-scale gamma                       */
#define PSSR_OP_COPY 8        /* channel-slice copy between NHWC buffers                    */
#define PSSR_OP_TAILSUM 9     /* 9-tap gather of the fused Reconstruction tail + *128+128   */
#define PSSR_OP_STEM 10       /* RDNet PatchifyStem: normalise + patch conv + LayerNorm2d   */
#define PSSR_OP_CAST8 11      /* 16-bit NHWC view * scale -> e5m2 NHWC (operand of PSSR_SEG_E5M2 segments) */
#define PSSR_OP_RESAMPLE 12   /* per-channel affine (+ReLU) / k x k max pool / bilinear enlargement of an NHWC view */
#define PSSR_OP_WINATTN 13    /* SwinIR shifted-window multi-head attention on an NHWC qkv map            */

/* One NHWC source view of an implicit-GEMM op. */
typedef struct {
  const void* base;      /* address of channel 0 of the view (16-byte aligned)             */
  int32_t channels;      /* valid channels of the view                                     */
  int32_t cstride;       /* elements between consecutive pixels (>= channels, % 8 == 0)    */
  int32_t H, W, B;       /* spatial size / batch of the source tensor                      */
  int32_t reserved;
} pssr_src_t;

/* A run of K blocks: for every tap (dy,dx) of the filter and every 64-channel block of
 * the source, one [128 pixels x 64] operand tile is fetched by TMA (out-of-image taps are
 * zero-filled by the TMA unit = the convolution's zero padding). */
typedef struct {
  int32_t src;           /* index into srcs[]                                              */
  int32_t taps;          /* 1: 1x1,  9: 3x3 pad 1,  4: 2x2 stride 2                        */
  int32_t cblocks;       /* 64-channel blocks per tap                                      */
  int32_t fmt;           /* PSSR_SEG_F16: 16-bit source x `weights`;  PSSR_SEG_E5M2: the source is an e5m2 NHWC tensor
                            (1-byte elements; channels / cstride in elements) multiplied with `weights8` by kind::f8f6f4
                            MMAs (K = 32, twice the 16-bit rate) into the same fp32 accumulator -- the low-order terms of
                            the compensated precision only have to be known to a few bits.  3x3 segments of layers whose
                            width is a multiple of 128 only.                                                          */
  int32_t dilation;      /* taps == 9 only: tap (dy,dx) reads the source at (y + dy*dilation, x + dx*dilation), zeros outside the
                            image = nn.Conv2d(padding="same", dilation=d) of ResBlockA (pssr/models/_blocks.py:56).  0 and 1 mean dense. */
} pssr_kseg_t;
#define PSSR_SEG_F16 0
#define PSSR_SEG_E5M2 1

typedef struct {
  pssr_src_t srcs[4];
  int32_t n_srcs;
  pssr_kseg_t segs[6];
  int32_t n_segs;
  const void* weights;   /* [n][k_total] 16-bit, K-major, K ordered seg -> tap -> cblock -> c */
  const float* bias;     /* [n] fp32 (BatchNorm shift and conv biases folded)               */
  int32_t n;             /* GEMM N = output channels incl. zero padding (multiple of 32)    */
  int32_t n_valid;       /* channels actually stored (<= n, multiple of 8)                  */
  int32_t Ho, Wo, B;     /* output pixel grid before pixel shuffle                          */
  void* out;             /* NHWC 16-bit [B][Ho*r][Wo*r][out_cstride], written at out_choff  */
  int32_t out_cstride;
  int32_t out_choff;
  int32_t shuffle;       /* r: F.pixel_shuffle factor applied by the epilogue (1 = none);   */
                         /* weights' N is ordered (i*r+j)*C' + c' so it is pure addressing  */
  int32_t act;           /* PSSR_ACT_*                                                      */
  const float* out_scale;/* optional per-channel multiplier applied after act (layer-scale) */
  float* out_f32;        /* optional fp32 NHWC copy of the output (same geometry) or NULL   */
  /* Fused Reconstruction tail (pssr/models/_blocks.py:15-18): when tail_z != NULL the epilogue does not
   * store the activation (out may be NULL).  Instead, per output pixel and per pixel-shuffle sub-position
   * s = i*r+j it reduces the C' = n/r^2 post-ReLU channels against the 3x3 tail weights in fp32:
   *     z[b][y][s*9 + t][x] = sum_c tail_weight[t][c] * act(acc[s*C' + c] + bias)        t = 0..8
   * (a 1x1 "per-tap" projection; PSSR_OP_TAILSUM then gathers the 9 shifted taps at HR resolution).
   * The r^2*C'-channel HR feature map -- 2 GB per 64-tile batch -- never exists in memory.            */
  const float* tail_weight; /* fp32 [9][C'] (single output channel)                               */
  float* tail_z;            /* fp32 [B][Ho][r*r*9][Wo]: plane rows of one LR row are contiguous   */
  /* tail_layout = PSSR_TAIL_WINDOW48 (r = 4, Wo % 128 == 0 only): the epilogue also pre-sums the projections of the 16
   * sub-positions of an LR pixel by the HR OUTPUT position they feed.  (s = (i', j'), tap (dy, dx)) feeds the output at
   * (oi, oj) = (i' - dy, j' - dx) in [-1, 4]^2 relative to the pixel's 4x4 HR block; sources with j' < 2 and j' >= 2 are
   * summed by different threads and stored separately (two 6 x 4 windows):
   *     z[b][y][e*24 + (oi+1)*4 + (oj+1-2e)][x],  e = j' / 2        -- 48 floats per LR pixel instead of 144.          */
  int32_t tail_layout;
  /* Compensated ("fp16c") precision, the mode that meets the 1e-2 max-abs bar of the network output: a 16-bit tensor t is
   * carried as the pair (hi = rn16(t), lo = rn16(t - hi)) where it feeds a sensitive layer, and the K segments of that layer
   * multiply hi*W_hi + lo*W_hi + hi*W_lo (the host packs W_hi / W_lo as ordinary K blocks; segments that read the same
   * source share its staged tile).
   *   out_lo     : optional second 16-bit NHWC output = rn16(y - rn16(y)) of the activated result y, same pixel grid as out
   *   tail_flags : PSSR_TAIL_COMP -- the fused tail projects hi and lo of relu(pre) on W_hi and W_lo of the tail weights
   *                (lo rides an e5m2 kind::f8f6f4 pass: it only has to be known to a few bits)                              */
  int32_t tail_flags;
  void* out_lo;
  int32_t out_lo_cstride;
  int32_t out_lo_choff;
  const void* weights8;  /* [n][k8_total] e5m2, K-major, K ordered like `weights` over the PSSR_SEG_E5M2 segments only */
  /* optional residual added in the epilogue before the activation: y = act(acc + bias + resid_scale * resid[pixel][n])
   * (16-bit NHWC on the output pixel grid, shuffle == 1).  The compensated ResBlock tail adds the low-order terms of its
   * 1x1 residual this way: a separate small GEMM computes them scaled by 2^10, so the main launch keeps its three source planes. */
  const void* resid;
  int32_t resid_cstride;
  int32_t resid_choff;
  float resid_scale;
  int32_t reserved3;
} pssr_conv_desc_t;
#define PSSR_TAIL_TAPS 0
#define PSSR_TAIL_WINDOW48 1
#define PSSR_TAIL_COMP 1

typedef struct {
  const void* x;         /* [B][C][H][W] float32 (0..255) or uint8 if x_u8                   */
  int32_t x_u8;
  int32_t B, C, H, W;
  const float* scale;    /* [C] BN scale  s  (1 if no norm)                                  */
  const float* shift;    /* [C] BN shift  t                                                   */
  void* im2col;          /* NHWC 16-bit [B][H][W][cols]: channel (c*9+tap) = norm(x) at the tap,
                            zero outside the image and for channels >= 9*C                   */
  void* im2col_lo;       /* optional: same layout, rn16(v - rn16(v)) of every value v of im2col (compensated precision) */
  float* xnorm_f32;      /* [B][C][H][W] fp32 normalised input (the final skip), may be NULL */
  int32_t cols;          /* channels per pixel of im2col: 64 (0 = 64) or 16 when 9*C <= 16 -- the consuming
                            convolutions' TMA boxes zero-fill channels >= cols, 4x less HBM traffic         */
  int32_t centre_only;   /* 1: no taps -- im2col is the normalised input itself, NHWC [B][H][W][cols] with cols a multiple
                            of 8 >= C (inputs with more than 7 channels: the first convolution is then an ordinary 3x3
                            segment over this tensor, whose TMA zero fill IS the post-normalisation padding)            */
} pssr_prep_desc_t;

typedef struct {
  const void* in; int32_t in_cstride; int32_t in_choff;
  void* out; int32_t out_cstride; int32_t out_choff;
  int32_t B, H, W, C;    /* input spatial size; output is H/2 x W/2                          */
} pssr_pool_desc_t;

typedef struct {
  const void* in;        /* NHWC 16-bit [B][H][W][cstride], C valid                          */
  int32_t cstride, C;
  int32_t B, H, W;
  const float* weight;   /* [Cout][3][3][C] fp32                                             */
  const float* bias;     /* [Cout]                                                           */
  int32_t Cout;
  float mul, add;        /* y = acc*mul + add  (resunet.py:95: 128, 128)                     */
  float* out_f32;        /* [B][Cout][H][W] fp32 or NULL                                     */
  uint8_t* out_u8;       /* [B][H][W] uint8 = trunc(clip(y,0,255)) of channel Cout/2
                            (predict.py:245-246 `_pred_array`) or NULL                       */
} pssr_tail_desc_t;

/* Second half of the fused tail: out[b][0][Y][X] = (bias + sum_t zHR[(Y+dy, X+dx)][t]) * mul + add with
 * zHR[(Y,X)][t] = z[b][Y/r][((Y%r)*r + X%r)*9 + t][X/r] and zero outside the HR image
 * (= Reconstruction.conv's zero padding), plus the fused `_pred_array` uint8 output.               */
typedef struct {
  const float* z;        /* [B][H][r*r*9][W], or [B][H][48][W] with layout PSSR_TAIL_WINDOW48  */
  int32_t B, H, W, r;    /* LR geometry and pixel-shuffle factor; output is [B][1][H*r][W*r]  */
  float bias, mul, add;
  int32_t layout;        /* PSSR_TAIL_TAPS / PSSR_TAIL_WINDOW48 (must match the producing conv) */
  float* out_f32;        /* [B][1][H*r][W*r] or NULL                                          */
  uint8_t* out_u8;       /* [B][H*r][W*r] or NULL                                             */
} pssr_tailsum_desc_t;

/* PSSR_OP_CAST8: out[b][y][x][c] = e5m2(in[b][y][x][c] * scale), round to nearest, saturating (C % 16 == 0). */
typedef struct {
  const void* in; int32_t in_cstride, in_choff, C;
  int32_t B, H, W;
  float scale;
  void* out; int32_t out_cstride, out_choff;
} pssr_cast8_desc_t;

/* PSSR_OP_RESAMPLE: the CUDA-core pieces of the atrous / PSP variants (pssr/models/_blocks.py:43-92) on 16-bit NHWC views.
 *   mode 0  out = act(in * scale[c] + shift[c])            the BatchNorm(eval) -> ReLU that PRECEDES each first convolution of a
 *                                                          ResBlockA branch (:52-54; it cannot fold into a conv across the ReLU)
 *   mode 1  out[b][y][x] = max over the k x k window       F.max_pool2d(x, kernel_size=k) (:85), output [B][H/k][W/k]
 *   mode 2  bilinear enlargement to Ho x Wo                F.interpolate(size=size, mode="bilinear") (:85): align_corners=False,
 *                                                          src = max((dst + 0.5) * in/out - 0.5, 0), fp32 lerp
 *   mode 3  out[c] = c < k ? in[in_choff + c] : 0          torch.chunk pieces whose width is not a multiple of 8 channels (in_choff and
 *                                                          in_cstride may be any value here), re-laid 8-aligned and zero-padded to C   */
typedef struct {
  const void* in; int32_t in_cstride, in_choff, C; int32_t B, H, W;
  int32_t mode, k, Ho, Wo, relu;              /* relu: 0 none, 1 ReLU, 2 LeakyReLU(0.01) (swinir.py:171) */
  const float* scale; const float* shift;     /* mode 0: [C] fp32 (NULL = identity)              */
  void* out; int32_t out_cstride, out_choff;
} pssr_resample_desc_t;

/* PSSR_OP_WINATTN: the attention of one SwinTransformerBlock (pssr/models/swinir.py:335-373, :563-592) between its qkv and proj
 * GEMMs.  qkv: NHWC 16-bit [B][H][W][cstride], channel = which*C + head*(C/heads) + e as nn.Linear(dim, 3*dim) leaves it.  The map is
 * shifted by -shift, cut into ws x ws windows; per window and head softmax((q*scale) k^T + bias + mask) v, where
 * biasT[head][j][i] = relative_position_bias_table[relative_position_index[i][j]][head] (fp32, TRANSPOSED) and the mask is -100 between
 * tokens of different shifted regions (calculate_mask, :320-341; absent when shift == 0).  Tokens return to their unshifted place in
 * out [B][H][W][out_cstride] at out_choff.  H, W multiples of ws; ws*ws <= 64; C/heads even and <= 32.                      */
typedef struct {
  const void* qkv; int32_t cstride, C, heads; int32_t B, H, W; int32_t ws, shift; float scale; int32_t reserved;
  const float* biasT;
  void* out; int32_t out_cstride, out_choff;
} pssr_winattn_desc_t;

/* ---- RDNet encoder ops (pssr/models/_rdnet.py) ---------------------------------------------- */
/* PSSR_OP_STEM: x/128-1 -> BatchNorm(eval) -> PatchifyStem conv (kernel = stride = patch, _rdnet.py:106-116)
 * -> LayerNorm2d over channels (timm, eps 1e-6).  Output NHWC 16-bit [B][H/patch][W/patch][.].        */
typedef struct {
  const void* x; int32_t x_u8; int32_t B, C, H, W;
  const float* in_scale; const float* in_shift;     /* [C] folded input BatchNorm                      */
  int32_t patch, Cout;
  const float* weight;   /* [Cout][C*patch*patch] fp32                                               */
  const float* bias; const float* ln_w; const float* ln_b; float eps; int32_t reserved;
  void* out; int32_t out_cstride, out_choff;
  void* out_lo;          /* compensated precision (optional): rn16(y - rn16(y)), same layout as `out`               */
} pssr_stem_desc_t;

/* PSSR_OP_LAYERNORM: LayerNorm2d over C of an NHWC 16-bit view (transition layers, _rdnet.py:57-58).  With
 * s2d = 2 the output is space-to-depth'ed: [B][H/2][W/2][(dy*2+dx)*C + c], which turns the following 2x2
 * stride-2 transition conv (_rdnet.py:59-62) into a 1x1 GEMM.                                           */
typedef struct {
  const void* in; int32_t in_cstride, in_choff, C; int32_t B, H, W; int32_t s2d;
  const float* w; const float* b; float eps; int32_t reserved;
  void* out; int32_t out_cstride, out_choff;
  const void* in_lo;     /* compensated precision (optional): the input is in + in_lo (same layout as `in`) ...       */
  void* out_lo;          /* ... and what the 16-bit rounding of the output dropped goes here (same layout as `out`)   */
} pssr_ln_desc_t;

/* PSSR_OP_DWCONV_LN: depthwise 7x7 conv (pad 3) + bias + LayerNorm2d (Block, _rdnet.py:181-183).     */
typedef struct {
  const void* in; int32_t in_cstride, in_choff, C; int32_t B, H, W; int32_t reserved;
  const float* dw_w;     /* [49][C] fp32                                                             */
  const float* dw_b; const float* ln_w; const float* ln_b; float eps; int32_t reserved2;
  void* out; int32_t out_cstride, out_choff;
  const void* in_lo;     /* compensated precision (optional), as in pssr_ln_desc_t; with out_lo set the pre-LayerNorm */
  void* out_lo;          /* values also travel as a hi + lo pair                                                     */
} pssr_dwln_desc_t;

/* PSSR_OP_ESE: EffectiveSEModule (timm) + layer-scale gamma (_rdnet.py:172-174,200-202):
 * out[b][y][x][choff+c] = in[b][y][x][c] * hardsigmoid(fc(mean_yx in[b]))[c] * gamma[c].               */
typedef struct {
  const void* in; int32_t in_cstride, C; int32_t B, H, W; int32_t reserved;
  const float* fc_w;     /* [C][C] fp32                                                              */
  const float* fc_b; const float* gamma;
  float* gate_ws;        /* [B][C] fp32 scratch                                                      */
  void* out; int32_t out_cstride, out_choff;
} pssr_ese_desc_t;

typedef struct {
  int32_t kind;          /* PSSR_OP_*                                                       */
  int32_t reserved;
  union {
    pssr_conv_desc_t conv;
    pssr_prep_desc_t prep;
    pssr_pool_desc_t pool;
    pssr_tail_desc_t tail;
    pssr_tailsum_desc_t tailsum;
    pssr_stem_desc_t stem;
    pssr_resample_desc_t resample;
    pssr_winattn_desc_t winattn;
    pssr_ln_desc_t ln;
    pssr_dwln_desc_t dwln;
    pssr_ese_desc_t ese;
    pssr_cast8_desc_t cast8;
    uint8_t pad[512];
  } u;
} pssr_op_t;

typedef struct pssr_plan pssr_plan_t;

/* Builds TMA descriptors for every op.  `dtype` is PSSR_DT_*.  The op list is copied. */
int pssr_plan_create(const pssr_op_t* ops, int32_t n_ops, int32_t dtype, pssr_plan_t** out);
/* Launches every op of the plan in order on `stream`. */
int pssr_plan_run(pssr_plan_t* plan, void* stream);
/* Launches ops [first, first+count) only (profiling / tests). */
int pssr_plan_run_range(pssr_plan_t* plan, int32_t first, int32_t count, void* stream);
int32_t pssr_plan_num_ops(const pssr_plan_t* plan);
void pssr_plan_destroy(pssr_plan_t* plan);

/* ------------------------------------------------------------------------------------
 * Family 3: overlap-weighted tile stitch  (pssr/util.py:116-137 `_patch_images`, :96-100).
 * tiles [n_stacks][n_rows*n_cols][T][T] uint8 -> sheets [n_stacks][n_rows*step+ov][n_cols*step+ov]
 * uint8 with step = T-ov; every output pixel gathers its <=4 contributors (no atomics),
 * sum / count in exact integer arithmetic == the reference's float64 sum/count truncated.
 * ------------------------------------------------------------------------------------ */
int pssr_stitch(const uint8_t* tiles, uint8_t* sheets, int32_t n_stacks, int32_t n_rows,
                int32_t n_cols, int32_t tile, int32_t overlap, int32_t margin, void* stream);

/* ------------------------------------------------------------------------------------
 * Family 4: scoring  (pssr/predict.py:193-203, skimage PSNR / SSIM, pssr/util.py:139-191).
 * ------------------------------------------------------------------------------------ */
/* Per image pair (uint8 [n][h][w]): exact integer reductions
 *   sums[i][0] = sum (a-b)^2  (int64)
 *   ssim[i]    = sum over the (h-6)x(w-6) interior of the 7x7-window SSIM map (double)
 * from which mse / pixel / psnr / ssim follow on the host (pssr2_b200/predict.py).
 * workspace: >= pssr_metric_workspace_bytes(n, h, w) bytes of device scratch (per-CTA partials). */
int64_t pssr_metric_workspace_bytes(int32_t n, int32_t h, int32_t w);
int pssr_metric_sums(const uint8_t* a, const uint8_t* b, int32_t n, int32_t h, int32_t w,
                     int64_t* sq_err, double* ssim_sum, void* workspace, void* stream);
/* normalize_preds (util.py:139-191) for uint8 pairs of equal shape: 256-bin histograms give
 * the exact percentiles/means; second pass applies the affine maps, clips and truncates.
 * workspace: >= pssr_normalize_workspace_bytes(n) bytes of device scratch. */
int64_t pssr_normalize_workspace_bytes(int32_t n);
int pssr_normalize_preds(const uint8_t* hr, const uint8_t* hr_hat, uint8_t* hr_out,
                         uint8_t* hr_hat_out, int32_t n, int32_t h, int32_t w, double pmin,
                         double pmax, void* workspace, void* stream);

/* normalize_preds with hr_hat at a LOWER resolution than hr (pssr/util.py:179: `resize(hr_hat_norm, hr_norm.shape)` feeds the covariance
 * only; `_collage_preds` normalises the low-resolution input against hr this way, pssr/predict.py:220-222).  The enlargement is
 * skimage.transform.resize(order 1, mode "reflect") = scipy.ndimage.zoom(order=1, mode="mirror", grid_mode=True), evaluated in double.
 * hr [n,h,w], hr_hat [n,hat_h,hat_w] uint8; outputs keep their own resolutions (either may be NULL).
 * workspace: pssr_normalize_resized_workspace_bytes(n) bytes of device memory, 16-byte aligned. */
int64_t pssr_normalize_resized_workspace_bytes(int32_t n);
int pssr_normalize_preds_resized(const uint8_t* hr, const uint8_t* hr_hat, uint8_t* hr_out, uint8_t* hr_hat_out, int32_t n, int32_t h, int32_t w,
                                 int32_t hat_h, int32_t hat_w, double pmin, double pmax, void* workspace, void* stream);

/* Noise-profile histogram of `approximate_crappifier`'s objective (pssr/train.py:366-380):
 * profile = float32(a) - float32(base); hist511 = np.histogram(profile, np.arange(-256, 256)) (int64 [511], last bin closed);
 * *sum = sum of the profile (double).  a_kind: 0 uint8, 1 float32, 2 float64.  Both outputs are device memory, zeroed here. */
int pssr_profile_hist(const void* a, int32_t a_kind, const uint8_t* base, int64_t n, int64_t* hist511, double* sum, void* stream);

/* ------------------------------------------------------------------------------------
 * File I/O edges: the reference reads sheets with tifffile.imread (pssr/data.py:566-571, :621-625) and
 * writes predictions / stitched sheets with tifffile.imwrite (pssr/predict.py:71, pssr/util.py:103).  These entry points
 * decode a grayscale 8 / 16-bit TIFF stack (uncompressed strips, one IFD per frame or a contiguous ImageJ hyperstack,
 * classic or BigTIFF, either byte order) straight into caller-owned -- normally pinned -- host memory, so the upload can
 * follow without a staging copy, and encode uint8 / uint16 stacks.  Host functions: no stream, callable from any thread.
 * ------------------------------------------------------------------------------------ */
/* native = 1: pssr_tiff_read can decode the file; 0: compressed / tiled / colour (the host mirror falls back to Pillow). */
int pssr_tiff_probe(const char* path, int32_t* frames, int32_t* h, int32_t* w, int32_t* bits, int32_t* native);
/* dst: [frames][h][w] in the file's bit depth, native byte order; dst_bytes >= frames*h*w*bits/8. */
int pssr_tiff_read(const char* path, void* dst, int64_t dst_bytes);
int pssr_tiff_write(const char* path, const void* src, int32_t frames, int32_t h, int32_t w, int32_t bits);

#ifdef __cplusplus
}
#endif
#endif /* PSSR_B200_H */
