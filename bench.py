#!/usr/bin/env python
"""bench.py -- HR megapixels/s of the PSSR2 test/predict hot path on B200 (BASELINE.json metric).

Main workload (N=1): BASELINE.json configs[1] -- ResUNet scale=4, batch 64 of 128->512 tiles, synthetic
16-bit EM-like images, random-init weights.  One step = one pass of the hot path over one batch:
fused crappify (HR uint16 tiles -> LR, Poisson + AdditiveGaussian, on-device Philox) -> ResUNet forward
(tcgen05 implicit-GEMM plan, fused `_pred_array`) -> PSNR/SSIM/MSE sums against the HR tiles.
With N > 1 (torchrun, one rank per GPU) every rank processes its own batch (tile-wise sharding, weak
scaling); the metric sums accumulate on the device and are all-reduced over NCCL once, inside the timed region.

  value : whole-job HR MP/s with the HR tiles already resident in HBM
  e2e   : the same through the public API (ImageDataset + predict_images) from pinned HOST buffers,
          H2D of the tiles and D2H of the uint8 predictions inside the timed region
  roofline / cpu_baseline : see DESIGN.md "Measurement"
  configs : bounded sub-records of the other BASELINE configs, measured in the same run
            "3" RDResUNet over 4096^2 sheets via SlidingDataset -> predict_sheets (device-resident stitch, sheets sharded)
            "4" crappifier-only GB/s (2048^2 -> 512^2 tiles, Poisson + AdditiveGaussian, uint16 and uint8)
            "5" ResUNet([5,1], scale 8) test_metrics over 5 x 2048^2 tiles
  --config {3,4,5} : that config alone as the main line (full-size sample)
  --impl reference : the reference's own CPU implementation (baseline/_ref: the unmodified pssr 2.4.0 wheel; its absent
                     third-party imports served by oracle/refshim.py) on the host cores, same metric / config.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TILE, SCALE, BATCH = 512, 4, 64
ALG_FLOPS_PER_TILE = 63.305e9          # ResUNet scale 4 conv FLOPs per 128^2 -> 512^2 tile
TAIL_FLOPS_PER_TILE = 0.302e9          # Reconstruction.conv: fused into the tail projections, excluded from the conv roofline
RD_FLOPS_PER_TILE = 107.15e9           # RDResUNet per 128^2 -> 512^2 tile
S8_FLOPS_PER_TILE = 512.06e9           # ResUNet([5,1], scale 8) per 5 x 256^2 -> 2048^2 tile
METRIC, UNIT = "HR megapixels/sec (ResUNet 4x)", "HR MP/s"


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["bf16_tflops"]), float(p["hbm_gbs"]), "measured"
    except Exception:
        return 1590.0, 6650.0, "fallback"


def _peak_sustained(burst_tf):
    """bf16_tflops_sustained of MEASURED_PEAKS.json (cuBLAS back to back for seconds under the power cap); the profiling
    recipe's stated fallback (1.4 PFLOP/s) when the file is absent."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["bf16_tflops_sustained"])
    except Exception:
        return 1400.0


# dram__bytes_read.sum + dram__bytes_write.sum summed over the conv launches of ONE step, from the ncu capture named below
CONV_DRAM_BYTES = {"fp16c": 4763400000, "fp16": 3876322304, "bf16": 3876322304}
CONV_DRAM_SOURCE = {"fp16c": "profiles/r02_final_launches_step_summary.txt (ncu launch list of `bench.py --kernels-only --no-sub`, 38 conv launches of "
                             "one step of the compensated plan: 3467.1 MB read + 1296.3 MB written)",
                    "fp16": "profiles/r01_launches_v3_summary.txt (37 conv launches of one step of the single-pass plan: 2823.8 MB read + 1052.6 MB written)"}
CONV_DRAM_SOURCE["bf16"] = CONV_DRAM_SOURCE["fp16"]
RECON_FLOPS_PER_TILE = 19.629e9
RECON_DRAM_BYTES = {"fp16c": 449149184, "fp16": 325600000, "bf16": 325600000}
RECON_DRAM_SOURCE = {"fp16c": "profiles/r02_final_v3_recon_pre.details.txt (ncu --set full: 280.2 MB read + 168.9 MB written)",
                     "fp16": "profiles/r01_launches_v3_summary.txt launch #42 (175.1 MB read + 150.5 MB written)"}
RECON_DRAM_SOURCE["bf16"] = RECON_DRAM_SOURCE["fp16"]


def _synthetic_tiles(n, seed, device, size=TILE, frames=None, dtype="u16"):
    """Microscopy-like tiles in a 0..255 range: smooth structure + shot noise; uint16 container (int16 view) or uint8."""
    import torch
    import torch.nn.functional as F
    g = torch.Generator(device=device).manual_seed(seed)
    c = 1 if frames is None else frames
    base = torch.rand(n, c, max(8, size // 21), max(8, size // 21), generator=g, device=device)
    base = F.interpolate(base, size=(size, size), mode="bicubic", align_corners=False).clamp(0, 1) * 160 + 20
    img = torch.poisson(base, generator=g).clamp(0, 255)
    img = img[:, 0] if frames is None else img
    return img.to(torch.int16 if dtype == "u16" else torch.uint8).contiguous()


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")] + [time.perf_counter()])

    def stop(self, t0=None, t1=None):
        """Samples taken inside the timed region [t0, t1] (host clock); all samples if the region was too short to hold one."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        if t0 is not None:
            inside = [r for r in self.rows if t0 <= r[-1] <= t1 + 0.02]
            if inside:
                self.rows = inside
        sm = sorted(int(r[1]) for r in self.rows if len(r) >= 8 and r[1].isdigit())
        mx = [int(r[2]) for r in self.rows if len(r) >= 8 and r[2].isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------- reference / CPU arm
def _host_threads():
    """All the host threads the CPU arm may use -- set explicitly, so that a torchrun launch (which exports OMP_NUM_THREADS=1)
    times the same thing as a plain launch."""
    import torch
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        pass
    torch.set_num_threads(n)
    return n


def _cpu_reference_steps(steps, warmup, tiles_per_step):
    """The reference's CPU path for BASELINE configs[1], sampled: per tile `_gen_pair` (square crop, Pillow BILINEAR downscale,
    MultiCrappifier(Poisson(), AdditiveGaussian()), round / clip: pssr/data.py:471-495), the fp32 ResUNet forward
    (pssr/models/resunet.py:65-96), `_pred_array` and PSNR / SSIM / MSE (pssr/predict.py:144-211) -- driven through the
    reference's OWN `test_metrics(model, dataset, device="cpu", norm=False)` when the unmodified package is importable
    (baseline/_ref or /root/reference; kind "reference"), else through the oracle port of the same calls (kind "port").
    Returns (HR MP/s, seconds per step, threads, kind)."""
    import numpy as np
    import torch
    threads = _host_threads()
    rng = np.random.default_rng(1234)
    hr_tiles = rng.poisson(100, (tiles_per_step, 1, TILE, TILE)).clip(0, 255).astype(np.uint16)
    kind = "port"
    try:
        from oracle.refshim import import_reference, reference_available
        if reference_available():
            import_reference()
            from pssr import crappifiers as RC, data as RD
            from pssr.models import ResUNet as RefResUNet
            from pssr.predict import test_metrics as ref_test_metrics
            kind = "reference"
    except Exception as e:          # noqa: BLE001 -- an import problem of the optional real reference falls back to the port
        sys.stderr.write(f"bench: reference package not usable ({e!r}); timing the oracle port\n")
        kind = "port"

    if kind == "reference":
        torch.manual_seed(0)
        model = RefResUNet().eval()
        crap = RC.MultiCrappifier(RC.Poisson(), RC.AdditiveGaussian())

        class Tiles:                 # the duck type test_metrics reads; items come from the reference's _gen_pair
            val_idx = list(range(tiles_per_step))
            crop_res, is_lr = TILE, False

            def __len__(self):
                return tiles_per_step

            def __getitem__(self, i):
                # (test_metrics evaluates dataset[0] on every iteration, pssr/predict.py:180: the same work per item)
                return RD._gen_pair(hr_tiles[i], TILE, SCALE, False, crap, None, None)

        def one_step():
            np.random.seed(0)
            import contextlib
            import io as _io
            with contextlib.redirect_stderr(_io.StringIO()):
                ref_test_metrics(model, Tiles(), device="cpu", norm=False)
    else:
        from oracle import pipeline as OP
        from oracle.models import resunet_forward
        from pssr2_b200.models import ResUNet
        torch.manual_seed(0)
        sd = {k: v.clone() for k, v in ResUNet().eval().state_dict().items()}

        def one_step():
            np.random.seed(0)
            for t in hr_tiles:
                lr0 = OP.resize_bilinear(t, TILE // SCALE, TILE // SCALE).astype(np.float32)
                stages = [("poisson", np.random.poisson(np.clip(lr0, 0, np.inf)), 1, 0), ("gaussian", np.random.normal(0, 13, lr0.shape))]
                hr, lr = OP.gen_pair(t, TILE, SCALE, stages)
                with torch.no_grad():
                    out = resunet_forward(sd, torch.as_tensor(lr)[None])
                a, b = OP.pred_array(hr[None]), OP.pred_array(out.numpy())
                OP.image_metrics(a[0], b[0])

    for _ in range(warmup):
        one_step()
    t0 = time.perf_counter()
    for _ in range(steps):
        one_step()
    dt = (time.perf_counter() - t0) / steps
    return tiles_per_step * TILE * TILE / dt / 1e6, dt, threads, kind


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    tiles = 4 if args.steps <= 20 else 1          # bounded sample: the whole run stays within a few minutes of CPU work
    warm = min(args.warmup, 1)
    mp, dt, threads, kind = _cpu_reference_steps(args.steps, warm, tiles)
    what = ("the unmodified reference package (pssr 2.4.0 wheel in baseline/_ref): its test_metrics / _gen_pair / ResUNet on torch CPU fp32; "
            "scikit-image's PSNR / SSIM are served by oracle/thirdparty.py (package absent)") if kind == "reference" else \
           "oracle port (torch fp32 + NumPy / Pillow-exact)"
    sample = f"{tiles} tiles/step x {args.steps} steps of the same workload on {threads} host threads: {what}"
    line = {"metric": METRIC, "value": round(mp, 4), "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": warm,
            "ms_per_step": round(dt * 1e3, 2), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "impl": "reference",
            "config": {"workload": "ResUNet scale=4, 128->512 tiles, crappify+forward+metrics (BASELINE configs[1] sampled)",
                       "tiles_per_step": tiles},
            "cpu_baseline": {"value": round(mp, 4), "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": round(mp, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------ shared helpers
class Ctx:
    pass


def _setup():
    import torch
    import torch.distributed as dist
    from pssr2_b200 import _lib
    from pssr2_b200 import dist as D
    c = Ctx()
    c.world = int(os.environ.get("WORLD_SIZE", "1"))
    c.rank = int(os.environ.get("RANK", "0"))
    c.local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(c.local)
    c.dev = torch.device("cuda", c.local)
    if c.world > 1:
        # NCCL announces its version on STDOUT when the first communicator is created; the contract is ONE JSON line on stdout,
        # so stdout points at stderr until the communicator exists
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            D.init_from_env("nccl")
            warm = torch.zeros(1, device=c.dev)
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    _lib.lib()   # fails loudly if the CUDA extension is missing
    return c


def _barrier(c):
    import torch
    import torch.distributed as dist
    if c.world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def _max_over_ranks(c, seconds):
    import torch
    import torch.distributed as dist
    t = torch.tensor([seconds], dtype=torch.float64, device=c.dev)
    if c.world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def _quiet():
    import contextlib
    import io
    return contextlib.redirect_stderr(io.StringIO())


def _ev_ms(fn, reps=10, warm=3):
    import torch
    for k in range(warm):
        fn(k)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for k in range(reps):
        fn(k)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def _plan_conv_ms(plan, reps=3):
    """Per-op device time with the ops executed IN SEQUENCE (realistic cache state): one event after every op."""
    import torch
    n_ops = len(plan.records)
    acc = [0.0] * n_ops
    for _ in range(reps):
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(n_ops + 1)]
        torch.cuda.synchronize()
        evs[0].record()
        for i in range(n_ops):
            plan.run(i, 1)
            evs[i + 1].record()
        torch.cuda.synchronize()
        for i in range(n_ops):
            acc[i] += evs[i].elapsed_time(evs[i + 1]) / reps
    return acc


# ------------------------------------------------------------------------------ config 3
def config3(c, sheets_per_rank=2, size=4096, precision="fp16c"):
    """BASELINE configs[2]: RDResUNet scale 4 over 4096^2 uint16 sheets, SlidingDataset(hr_res=512, overlap=128) -> 100 tiles per
    sheet -> stitched 3968^2 sheets.  Sheets are sharded across ranks (weak scaling: `sheets_per_rank` each, ingested from pinned
    host memory); every rank predicts and stitches its own sheets on the device; rank 0 receives the stitched sheets over NCCL."""
    import numpy as np
    import torch
    from pssr2_b200 import dist as D, ops
    from pssr2_b200.crappifiers import AdditiveGaussian, MultiCrappifier, Poisson
    from pssr2_b200.data import SlidingDataset
    from pssr2_b200.models import RDResUNet
    from pssr2_b200.predict import predict_sheets
    torch.manual_seed(0)
    model = RDResUNet().eval()
    model.precision = precision
    model = model.to(c.dev)
    n_sheets = sheets_per_rank * c.world
    bsz = 50
    # every rank describes the whole dataset; a sheet's pixels exist only on the rank that owns it (pinned host memory), the
    # others hold a shape-only placeholder that sheet-aligned sharding never touches
    srcs = {}
    for s in range(n_sheets):
        if s // sheets_per_rank == c.rank:
            srcs[f"sheet{s}"] = _synthetic_tiles(1, 4321 + s, c.dev, size=size)[0:1].cpu().pin_memory()    # [1, size, size] uint16 container
        else:
            srcs[f"sheet{s}"] = _Placeholder((1, size, size))
    crap = MultiCrappifier(Poisson(), AdditiveGaussian())

    def make_ds():
        return SlidingDataset(dict(srcs), hr_res=512, lr_scale=4, overlap=128, val_split=1, crappifier=crap, device=c.dev, preload=False)

    with _quiet(), torch.no_grad():
        predict_sheets(model, make_ds(), device=str(c.dev), batch_size=bsz, margin=32)      # warm-up: plan build, allocator
        _barrier(c)
        t0 = time.perf_counter()
        out = predict_sheets(model, make_ds(), device=str(c.dev), batch_size=bsz, margin=32)
        torch.cuda.synchronize()
        dt = _max_over_ranks(c, time.perf_counter() - t0)
    if c.rank == 0:
        assert all(o is not None and o.shape == (1, 3968, 3968) for o in out), [None if o is None else o.shape for o in out]
    # forward-only time of the same plan (100 tiles = two batches of 50), and the stitch alone
    st = next(iter(model._plans.values()))
    fwd_ms = _ev_ms(lambda k: st["plan"].run(), reps=5, warm=2) * 2
    tiles = torch.randint(0, 255, (100, 512, 512), dtype=torch.uint8, device=c.dev)
    stitch_ms = _ev_ms(lambda k: ops.stitch(tiles, 10, 10, 128, 32), reps=10)
    stitch_bytes = 100 * 512 * 512 + 3968 * 3968
    _, hbm, _ = _peaks()
    per_sheet = dt / sheets_per_rank
    acc = _plan_conv_ms(st["plan"], reps=2)
    conv_ms = sum(t for t, (kind, _) in zip(acc, st["plan"].records) if kind == "conv") * 2
    peak_tf, _, _ = _peaks()
    return {"workload": f"RDResUNet scale=4, {n_sheets} synthetic uint16 sheets {size}^2 ({sheets_per_rank} per GPU), SlidingDataset(hr_res=512, overlap=128) "
                        "-> 100 tiles/sheet -> predict_sheets (crappify Poisson+AdditiveGaussian, forward, device-resident stitch margin 32) "
                        "-> 3968^2 uint8 sheets on the host of rank 0",
            "precision": precision, "sheets": n_sheets, "n_gpus": c.world, "seconds": round(dt, 4),
            "sheets_per_s": round(n_sheets / dt, 3), "stitched_mp_per_s": round(n_sheets * 3968 * 3968 / dt / 1e6, 1),
            "hr_mp_per_s_tiles": round(n_sheets * 100 * 512 * 512 / dt / 1e6, 1),
            "ms_per_sheet_e2e": round(per_sheet * 1e3, 2), "ms_per_sheet_forward": round(fwd_ms, 2),
            "e2e_over_forward": round(per_sheet * 1e3 / fwd_ms, 3),
            "forward_tflops_algorithmic": round(RD_FLOPS_PER_TILE * 100 / (fwd_ms * 1e-3) / 1e12, 1),
            "conv_tflops_algorithmic": round(RD_FLOPS_PER_TILE * 100 / (conv_ms * 1e-3) / 1e12, 1),
            "conv_frac_of_burst": round(RD_FLOPS_PER_TILE * 100 / (conv_ms * 1e-3) / 1e12 / peak_tf, 4),
            "stitch": {"ms": round(stitch_ms, 4), "algorithmic_bytes": stitch_bytes, "achieved_gbs": round(stitch_bytes / (stitch_ms * 1e-3) / 1e9, 1),
                       "frac_of_hbm_peak": round(stitch_bytes / (stitch_ms * 1e-3) / 1e9 / hbm, 4)},
            "h2d_bytes_per_sheet": size * size * 2, "d2h_bytes_per_sheet": 3968 * 3968}


class _Placeholder:
    """Shape / dtype of a sheet another rank owns (sheet-aligned sharding never reads it on this rank)."""

    path = "<sheet of another rank>"

    def __init__(self, shape):
        import numpy as np
        self.shape, self.dtype = tuple(shape), np.dtype(np.uint16)

    def read_pinned(self):
        raise RuntimeError("this rank does not own the sheet")

    def prefetch(self):
        pass

    def ready(self):
        return False


# ------------------------------------------------------------------------------ config 4
def config4(c, tiles_total=1024, distinct=64):
    """BASELINE configs[3]: crappifier-only throughput, Poisson + AdditiveGaussian downscale-crappify of 1024 tiles 2048^2 -> 512^2.
    `distinct` resident tiles (64 x 8.4 MB = 537 MB of uint16 > L2) are cycled; algorithmic bytes per tile = HR read + fp32 LR
    written (9.44 MB uint16 / 5.24 MB uint8)."""
    import torch
    from pssr2_b200 import ops
    from pssr2_b200.crappifiers import AdditiveGaussian, MultiCrappifier, Poisson
    _, hbm, src = _peaks()
    specs = MultiCrappifier(Poisson(), AdditiveGaussian()).noise_specs()
    out = {"workload": f"{tiles_total} tiles 2048^2 -> 512^2 per GPU, Poisson + AdditiveGaussian (Philox), {distinct} distinct resident tiles cycled "
                       f"({distinct * 2048 * 2048 * 2 / 1e6:.0f} MB uint16 > L2)", "peak_gbs": hbm, "peak_source": src, "n_gpus": c.world}
    chunk = 32
    for dt in ("u16", "u8"):
        pool = [_synthetic_tiles(chunk, 99 + i, c.dev, size=2048, dtype=dt) for i in range(distinct // chunk)]
        tables = [ops.TileTable([p], [0] * chunk, list(range(chunk)), [0] * chunk, [0] * chunk, [2048] * chunk, [2048] * chunk) for p in pool]
        eb = 2 if dt == "u16" else 1
        per_tile = 2048 * 2048 * eb + 512 * 512 * 4
        for label, st in (("noise", specs), ("resample_only", None)):
            launches = tiles_total // chunk

            def run(k, st=st):
                for j in range(launches):
                    ops.crappify(tables[j % len(tables)], 2048, 4, st, clip_between=True, seed=k, tile_index0=j * chunk)
            ms = _ev_ms(run, reps=3, warm=1)
            sec = _max_over_ranks(c, ms * 1e-3)
            gbs = per_tile * tiles_total / sec / 1e9
            out[f"{dt}_{label}"] = {"ms_per_1024_tiles": round(sec * 1e3, 3), "algorithmic_bytes_per_tile": per_tile,
                                    "achieved_gbs_per_gpu": round(gbs, 1), "frac_of_hbm_peak": round(gbs / hbm, 4),
                                    "tera_hr_px_per_s_all_gpus": round(c.world * tiles_total * 2048 * 2048 / sec / 1e12, 3)}
        del pool, tables
    return out


# ------------------------------------------------------------------------------ config 5
def config5(c, tiles_per_rank=8, precision="fp16c"):
    """BASELINE configs[4]: ResUNet(channels=[5,1], scale=8), 5-frame 2048^2 uint16 HR stacks -> 5 x 256^2 LR -> 2048^2, through
    test_metrics (crappify, forward, normalize_preds, PSNR / SSIM / MSE), tiles sharded across ranks (weak scaling)."""
    import torch
    from pssr2_b200.crappifiers import AdditiveGaussian, MultiCrappifier, Poisson
    from pssr2_b200.data import ImageDataset
    from pssr2_b200.models import ResUNet
    from pssr2_b200.predict import test_metrics
    torch.manual_seed(0)
    model = ResUNet(channels=[5, 1], scale=8).eval()
    model.precision = precision
    model = model.to(c.dev)
    n = tiles_per_rank * c.world
    stacks = [_synthetic_tiles(1, 777 + i, c.dev, size=2048, frames=5)[0].cpu().pin_memory() for i in range(tiles_per_rank)]
    stacks = stacks * c.world      # the dataset is sharded by item ranges: every rank reads its own block of (identical) stacks
    crap = MultiCrappifier(Poisson(), AdditiveGaussian())
    bsz = 4

    def run():
        ds = ImageDataset(stacks, hr_res=2048, lr_scale=8, n_frames=[5, 1], val_split=1, crappifier=crap, device=c.dev)
        return test_metrics(model, ds, device=str(c.dev), norm=True, item0_quirk=False, batch_size=bsz)

    # the run is short (16 tiles, ~18 ms) and bound by host->device copies, which other tenants of the box's host disturb: three
    # timed runs, the median is reported (all of them are in `seconds_all`)
    dts = []
    with _quiet(), torch.no_grad():
        run()
        for _ in range(3):
            _barrier(c)
            t0 = time.perf_counter()
            m = run()
            torch.cuda.synchronize()
            dts.append(_max_over_ranks(c, time.perf_counter() - t0))
    dt = sorted(dts)[1]
    st = next(iter(model._plans.values()))
    acc = _plan_conv_ms(st["plan"], reps=2)
    conv_ms = sum(t for t, (kind, _) in zip(acc, st["plan"].records) if kind == "conv")
    fwd_ms = _ev_ms(lambda k: st["plan"].run(), reps=5, warm=2)
    peak_tf, _, _ = _peaks()
    return {"workload": f"ResUNet(channels=[5,1], scale=8), {n} synthetic 5-frame 2048^2 uint16 stacks ({tiles_per_rank} per GPU): test_metrics(norm=True) = "
                        "crappify (Poisson+AdditiveGaussian) + forward + normalize_preds + PSNR/SSIM/MSE, batch 4, from pinned host stacks",
            "precision": precision, "tiles": n, "n_gpus": c.world, "seconds": round(dt, 4), "seconds_all": [round(t, 4) for t in dts],
            "hr_mp_per_s": round(n * 2048 * 2048 / dt / 1e6, 1), "ms_per_tile_e2e": round(dt / tiles_per_rank * 1e3, 2),
            "ms_per_tile_forward": round(fwd_ms / bsz, 3),
            "h2d_bytes_per_tile": 5 * 2048 * 2048 * 2,
            "h2d_gbs_all_gpus": round(n * 5 * 2048 * 2048 * 2 / dt / 1e9, 1),     # what bounds this config on several GPUs (shared host links)
            "forward_tflops_algorithmic": round(S8_FLOPS_PER_TILE * bsz / (fwd_ms * 1e-3) / 1e12, 1),
            "conv_tflops_algorithmic": round(S8_FLOPS_PER_TILE * bsz / (conv_ms * 1e-3) / 1e12, 1),
            "conv_frac_of_burst": round(S8_FLOPS_PER_TILE * bsz / (conv_ms * 1e-3) / 1e12 / peak_tf, 4),
            "metrics": {k: round(float(v), 5) for k, v in m.items()}}


# -------------------------------------------------------------------------------- config 2 (main line)
def run_ours(args):
    import torch
    import torch.distributed as dist
    from pssr2_b200 import _lib, ops
    from pssr2_b200.crappifiers import AdditiveGaussian, MultiCrappifier, Poisson
    from pssr2_b200.data import ImageDataset
    from pssr2_b200.models import ResUNet
    from pssr2_b200.predict import predict_images

    c = _setup()
    world, rank, dev = c.world, c.rank, c.dev
    if args.config in (3, 4, 5):
        with torch.no_grad():
            rec = {3: lambda: config3(c, sheets_per_rank=4, precision=args.precision), 4: lambda: config4(c),
                   5: lambda: config5(c, tiles_per_rank=16, precision=args.precision)}[args.config]()
        if rank == 0:
            print(json.dumps({"metric": f"BASELINE configs[{args.config - 1}]", "n_gpus": world, "data": "synthetic", "config": rec,
                              "note": "full-size sample of one secondary config; the driver's bench line is the default run"}))
        _barrier(c)
        return

    torch.manual_seed(0)
    model = ResUNet().eval()
    model.precision = args.precision
    model = model.to(dev)
    crap = MultiCrappifier(Poisson(), AdditiveGaussian())
    NB = 8   # resident input batches rotated between steps: 8 x 33.5 MB > 126 MB L2
    batches = [_synthetic_tiles(BATCH, 1234 + rank * 100 + i, dev) for i in range(NB)]
    tables = [ops.TileTable([b], [0] * BATCH, list(range(BATCH)), [0] * BATCH, [0] * BATCH, [TILE] * BATCH, [TILE] * BATCH) for b in batches]
    sums = torch.zeros(3, dtype=torch.float64, device=dev)      # [sum sq err, sum of the SSIM maps, images], accumulated on the device
    specs = crap.noise_specs()                                  # Poisson(i=1) + AdditiveGaussian(sigma=13), resolved once

    def step(i):
        lr, _, hr8 = ops.crappify(tables[i % NB], TILE, SCALE, specs, clip_between=True, seed=i, tile_index0=(rank * 1000003 + i) * BATCH,
                                  want_hr_u8=True)
        _, out8 = model.forward_u8(lr)
        sq, ss = ops.metric_sums(hr8[:, 0], out8[:, 0])
        sums[0] += sq.sum()
        sums[1] += ss.sum()
        sums[2] += BATCH

    for i in range(max(args.warmup, 3)):
        step(i)
    sums.zero_()
    _barrier(c)
    sampler = ClockSampler(c.local)
    if rank == 0 and not os.environ.get("PSSR_NO_CLOCK_SAMPLER"):
        sampler.start()
    launches0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    _barrier(c)
    if args.kernels_only:
        torch.cuda.profiler.start()     # ncu --profile-from-start off: the launch list holds exactly the timed steps
    t_host0 = time.perf_counter()
    e0.record()
    for i in range(args.steps):
        step(100 + i)
    if world > 1:
        dist.all_reduce(sums)           # the path's only collective here: ONE all-reduce of the metric sums, not one per step
    e1.record()
    _barrier(c)
    t_host1 = time.perf_counter()
    if args.kernels_only:
        torch.cuda.profiler.stop()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    launches = _lib.launch_count() - launches0
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_step = float(ms) / args.steps
    clocks = sampler.stop(t_host0, t_host1) if rank == 0 else None
    value = world * BATCH * TILE * TILE / (ms_step * 1e-3) / 1e6

    if args.kernels_only:
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": round(value, 1), "unit": UNIT, "ms_per_step": round(ms_step, 3),
                              "gpu_launches": int(launches), "precision": args.precision,
                              "note": "--kernels-only run (profiling), not a bench line"}))
        return

    # ---- e2e through the public API from pinned host buffers ---------------------------------
    # ONE predict_images call over a dataset of `e2e_steps` batches (one pinned host stack of 64 tiles per step): the timed
    # region holds the dataset construction (host -> device upload of every stack), crappify, forward, and the device -> host
    # read of every uint8 prediction; uploads and read-backs overlap the kernels of neighbouring batches.
    e2e_steps = max(2, min(args.steps, 50))      # 50 x 16.8 MB of pinned predictions
    host = [b.cpu().pin_memory() for b in batches[:2]]
    stacks = [host[i % 2] for i in range(e2e_steps)]

    def e2e_run():
        ds = ImageDataset(list(stacks), hr_res=TILE, lr_scale=SCALE, crappifier=crap, n_frames=1, val_split=1, device=dev)
        ds.rank_local = True          # weak scaling: every rank ingests and predicts its own stacks, nothing is gathered
        return predict_images(model, ds, device=str(dev), batch_size=BATCH, out_dir=None)

    import contextlib
    import io
    with contextlib.redirect_stderr(io.StringIO()), contextlib.redirect_stdout(io.StringIO()):
        preds = e2e_run()
        del preds
        # three timed calls (each: max over ranks), the MEDIAN is reported and all three are listed: the call is host-driven and
        # the pool's boxes are shared VMs (single calls scatter by ~5 %)
        e2e_all = []
        for _ in range(3):
            _barrier(c)
            t0 = time.perf_counter()
            preds = e2e_run()
            torch.cuda.synchronize()
            e2e_dt = torch.tensor([(time.perf_counter() - t0) / e2e_steps], dtype=torch.float64, device=dev)
            assert len(preds) == e2e_steps * BATCH
            del preds
            if world > 1:
                dist.all_reduce(e2e_dt, op=dist.ReduceOp.MAX)
            e2e_all.append(float(e2e_dt))
    e2e_dt = sorted(e2e_all)[1]
    e2e_value = world * BATCH * TILE * TILE / float(e2e_dt) / 1e6
    h2d = BATCH * TILE * TILE * 2
    d2h = BATCH * TILE * TILE

    # ---- the other BASELINE configs, bounded, in the same run (all ranks take part: they hold collectives) -------------
    subs = {}
    if not args.no_sub:
        with torch.no_grad():
            for key, fn in (("3", lambda: config3(c, sheets_per_rank=2, precision=args.precision)), ("4", lambda: config4(c, tiles_total=512)),
                            ("5", lambda: config5(c, tiles_per_rank=16, precision=args.precision))):
                try:
                    subs[key] = fn()
                except Exception as e:      # noqa: BLE001 -- a failing secondary config must not take the main line down
                    subs[key] = {"error": repr(e)[:300]}
                torch.cuda.empty_cache()

    if rank != 0:
        if world > 1:
            dist.barrier()
        return

    # ---- roofline of the dominant kernel (the tcgen05 implicit-GEMM conv), measured live -----------
    st = next(v for k, v in model._plans.items() if k[0][0] == BATCH)
    plan = st["plan"]
    acc = _plan_conv_ms(plan)
    conv_ms = sum(t for t, (kind, _) in zip(acc, plan.records) if kind == "conv")
    other_ms = sum(t for t, (kind, _) in zip(acc, plan.records) if kind != "conv")
    n_conv = sum(1 for kind, _ in plan.records if kind == "conv")
    peak_tf, peak_hbm, peak_src = _peaks()
    peak_sus = _peak_sustained(peak_tf)
    recon_ms = max((t for t, (kind, r) in zip(acc, plan.records) if kind == "conv" and r.get("tail_z") is not None), default=None)
    conv_flops = (ALG_FLOPS_PER_TILE - TAIL_FLOPS_PER_TILE) * BATCH
    issued = sum(r["issued_flops"] for kind, r in plan.records if kind == "conv")
    achieved = conv_flops / (conv_ms * 1e-3) / 1e12
    # The conv launches are timed inside the step sequence (tens of milliseconds of back-to-back tensor work under the 1 kW
    # cap), so the denominator is the SUSTAINED cuBLAS figure of MEASURED_PEAKS.json; the burst figure is reported beside it.
    roofline = {"bound": "tensor",
                "kernel": f"conv_v3_kernel (tcgen05 cta_group::2 implicit GEMM: rows mode at 128^2, cols mode below): all {n_conv} conv launches of the step",
                "achieved": round(achieved, 1), "peak": peak_sus, "unit": "TFLOP/s", "frac": round(achieved / peak_sus, 4),
                "traffic": CONV_DRAM_BYTES[args.precision], "traffic_source": CONV_DRAM_SOURCE[args.precision],
                "peak_source": peak_src + " (bf16_tflops_sustained)", "peak_burst": peak_tf, "frac_of_burst": round(achieved / peak_tf, 4),
                "algorithmic_flops_per_step": conv_flops, "issued_flops_per_step": issued,
                "issued_tflops": round(issued / (conv_ms * 1e-3) / 1e12, 1),
                "issued_note": "issued = K-padded MMA work incl. the hi/lo compensation segments of the fp16c plan (e5m2 segments counted at "
                               "their K, they run at twice the 16-bit rate); the judged figure uses the ALGORITHMIC flops",
                "launches_per_step": n_conv, "kernel_ms_per_step": round(conv_ms, 3), "other_net_ms_per_step": round(other_ms, 3),
                "dominant_launch": None if not recon_ms else {
                    "kernel": "conv_v3_kernel<T=1,G=1,TAIL,PAIR,ROWS>: Reconstruction.pre 65->1024 @128^2 + tensor-core tail projection",
                    "ms": round(recon_ms, 3), "achieved": round(RECON_FLOPS_PER_TILE * BATCH / (recon_ms * 1e-3) / 1e12, 1),
                    "peak": peak_sus, "frac": round(RECON_FLOPS_PER_TILE * BATCH / (recon_ms * 1e-3) / 1e12 / peak_sus, 4),
                    "frac_of_burst": round(RECON_FLOPS_PER_TILE * BATCH / (recon_ms * 1e-3) / 1e12 / peak_tf, 4),
                    "traffic": RECON_DRAM_BYTES[args.precision], "traffic_source": RECON_DRAM_SOURCE[args.precision]},
                "step_frac_of_peak": round(ALG_FLOPS_PER_TILE * BATCH / (ms_step * 1e-3) / 1e12 / peak_sus, 4)}

    # ---- the operand modes side by side (forward only, same weights): what the 1e-2 max-abs bar costs ---------
    modes = {}
    x_probe = torch.randint(0, 256, (BATCH, 1, TILE // SCALE, TILE // SCALE), device=dev).float()
    for prec, err in (("fp16c", "max-abs vs fp32 reference 5.7e-3 .. 7.4e-3 (tests/test_gpu_net.py, asserted <= 1e-2)"),
                      ("fp16", "max-abs 1.4e-2 .. 1.7e-2 (single pass; misses the 1e-2 bar)")):
        model.precision = prec
        stp, _ = model._state(x_probe)
        fwd = _ev_ms(lambda k: stp["plan"].run(), reps=10)
        modes[prec] = {"forward_ms": round(fwd, 3), "algorithmic_tflops": round(ALG_FLOPS_PER_TILE * BATCH / (fwd * 1e-3) / 1e12, 1),
                       "frac_of_burst": round(ALG_FLOPS_PER_TILE * BATCH / (fwd * 1e-3) / 1e12 / peak_tf, 4), "network_parity": err}
    model.precision = args.precision

    # ---- the HBM-side kernel families (crappify / tail gather / metrics / stitch), timed alone with CUDA events ---------
    # algorithmic bytes: DESIGN.md section 3 (every input byte read once + every output byte written once); peak = measured copy
    lr_b, _, hr8_b = ops.crappify(tables[0], TILE, SCALE, specs, clip_between=True, seed=1, want_hr_u8=True)
    _, out8_b = model.forward_u8(lr_b)
    out8_b = out8_b.clone()
    n_px = BATCH * TILE * TILE
    lr_px = n_px // (SCALE * SCALE)
    t_crap = _ev_ms(lambda k=0: ops.crappify(tables[k % NB], TILE, SCALE, specs, clip_between=True, seed=k, want_hr_u8=True))
    t_met = _ev_ms(lambda k=0: ops.metric_sums(hr8_b[:, 0], out8_b[:, 0]))
    tiles8 = out8_b[:, 0].repeat(8, 1, 1)        # 8 sheets of 8 x 8 tiles: long enough that the launch latency does not dominate
    t_stitch = _ev_ms(lambda k=0: ops.stitch(tiles8, 8, 8, 128, 32))
    tail_ms = sum(t for t, (kind, _) in zip(acc, plan.records) if kind == "tailsum")
    zbytes = sum(r["z"].numel() * 4 for kind, r in plan.records if kind == "tailsum")

    def hbm(ms, nbytes, what):
        return {"ms": round(ms, 4), "algorithmic_bytes": int(nbytes), "achieved_gbs": round(nbytes / (ms * 1e-3) / 1e9, 1),
                "frac_of_hbm_peak": round(nbytes / (ms * 1e-3) / 1e9 / peak_hbm, 4), "what": what}

    hbm_kernels = {
        "peak_gbs": peak_hbm, "peak_source": peak_src,
        "crappify": hbm(t_crap, n_px * 2 + lr_px * 4 + n_px, "64 uint16 HR tiles 512^2 read, float32 LR + uint8 HR written; Poisson+Gaussian Philox noise"),
        "tailsum": hbm(tail_ms, zbytes + n_px * 5, "window sums read once, fp32 + uint8 prediction written"),
        "metric_sums": hbm(t_met, n_px * 2, "two uint8 images read; SSIM window sums + SSE"),
        "stitch": hbm(t_stitch, 8 * (n_px + (8 * 384 + 128) ** 2), "8 x 64 uint8 tiles 512^2 (8x8 grids, overlap 128, margin 32) -> 8 sheets of 3200^2"),
    }

    # ---- CPU baseline: the reference (or its port) on this box's host cores, bounded sample ------------------------
    if world == 1:
        cpu_mp, cpu_dt, threads, kind = _cpu_reference_steps(3, 1, 8)
        cpu = {"value": round(cpu_mp, 4), "unit": UNIT, "cores": threads, "kind": kind,
               "sample": "3 steps x 8 tiles of the same workload after 1 warm-up step (%s: _gen_pair + fp32 forward + _pred_array + PSNR/SSIM/MSE), "
                         "%.1f s of CPU work" % ("the reference's own test_metrics from baseline/_ref" if kind == "reference" else "oracle port", 3 * cpu_dt)}
    else:
        cpu = None      # reported at N = 1 only

    line = {"metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": round(ms_step, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": {"fp16c": "fp16 operands with hi+lo compensation (+ e5m2 low-order terms) / f32 accumulate",
                                           "fp16": "fp16 operands / f32 accumulate", "bf16": "bf16 operands / f32 accumulate"}[args.precision],
            "data": "synthetic",
            "config": {"workload": "ResUNet scale=4, batch 64 of 128->512 tiles per GPU, synthetic 16-bit EM tiles: crappify "
                                   "(Poisson+AdditiveGaussian) + forward + PSNR/SSIM/MSE sums (BASELINE configs[1])",
                       "global_batch": BATCH * world, "tile": f"{TILE // SCALE}->{TILE}", "parallelism": f"tile-sharded x{world}",
                       "precision": args.precision + (" (default: meets the 1e-2 max-abs network tolerance; bf16 as named by the north star reaches "
                                                      "1.2e-1, single-pass fp16 1.5e-2 -- see precision_modes)" if args.precision == "fp16c" else ""),
                       "l2": f"inputs rotate over {NB} resident batches ({NB * h2d / 1e6:.0f} MB) and each step streams >4 GB of "
                             "activations, both > 126 MB L2"},
            "e2e": {"value": round(e2e_value, 1), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step_all_calls": [round(t * 1e3, 3) for t in e2e_all], "steps_per_call": e2e_steps,
                    "api": "ImageDataset(pinned host stacks, one per step) + one predict_images(batch_size=64, out_dir=None) call over all steps; "
                           "three calls back to back, median reported (the board heats towards its 1 kW cap during them: a call after a "
                           "pause runs ~4 % faster than the third one)"
                           + (" per rank (rank_local datasets: no gather)" if world > 1 else "")},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "precision_modes": modes, "hbm_kernels": hbm_kernels,
            "cpu_baseline": cpu, "configs": subs}
    s = [float(v) for v in sums.cpu()]
    mse = s[0] / max(s[2], 1) / (TILE * TILE)
    line["metric_check"] = {"images": int(s[2]), "mean_mse_255": round(mse, 3),
                            "psnr_of_mean_mse_db": round(10 * math.log10(255.0 ** 2 / mse), 3) if mse > 0 else None,
                            "mean_ssim": round(s[1] / max(s[2], 1) / ((TILE - 6) ** 2), 5)}
    print(json.dumps(line))
    if world > 1:
        dist.barrier()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="fp16c", choices=["fp16c", "fp16", "bf16"])
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4, 5], help="BASELINE config (1-based); 2 = the bench line with sub-records of 3-5")
    ap.add_argument("--no-sub", action="store_true", help="skip the sub-records of configs 3-5")
    ap.add_argument("--kernels-only", action="store_true", help="timed region only (for ncu launch lists): skip e2e / roofline / CPU legs")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
