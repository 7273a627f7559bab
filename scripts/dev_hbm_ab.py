"""Developer A/B timing of the HBM-side kernels (not a bench): metric sums, stitch (band vs per-row), crappify (3 vs 4 CTAs/SM)."""
import os, subprocess, sys
sys.path.insert(0, ".")
if len(sys.argv) > 1:
    import torch
    from pssr2_b200 import ops
    from pssr2_b200.crappifiers import AdditiveGaussian, MultiCrappifier, Poisson
    def ev(fn, reps=20):
        for _ in range(3): fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record()
        for _ in range(reps): fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / reps
    x = torch.randint(0, 255, (64, 512, 512), dtype=torch.uint8, device="cuda")
    y = (x.int() + torch.randint(-9, 9, x.shape, device="cuda")).clamp(0, 255).to(torch.uint8)
    t = ev(lambda: ops.metric_sums(x, y)); print(f"metric_sums 64x512^2: {t*1e3:.1f} us  {2*x.numel()/t/1e6:.0f} GB/s")
    t8 = x.repeat(8, 1, 1)
    t = ev(lambda: ops.stitch(t8, 8, 8, 128, 32)); print(f"stitch 8 sheets of 8x8 tiles: {t*1e3:.1f} us  {8*(64*512*512+3200*3200)/t/1e6:.0f} GB/s")
    t100 = torch.randint(0, 255, (100, 512, 512), dtype=torch.uint8, device="cuda")
    t = ev(lambda: ops.stitch(t100, 10, 10, 128, 32)); print(f"stitch 10x10 -> 3968^2: {t*1e3:.1f} us  {(100*512*512+3968*3968)/t/1e6:.0f} GB/s")
    specs = MultiCrappifier(Poisson(), AdditiveGaussian()).noise_specs()
    for dt in (torch.int16, torch.uint8):
        tiles = torch.randint(0, 255, (32, 2048, 2048), device="cuda").to(dt)
        tab = ops.TileTable([tiles], [0] * 32, list(range(32)), [0] * 32, [0] * 32, [2048] * 32, [2048] * 32)
        for label, st in (("noise", specs), ("plain", None)):
            t = ev(lambda: ops.crappify(tab, 2048, 4, st, clip_between=True), reps=5)
            nb = 32 * (2048 * 2048 * tiles.element_size() + 512 * 512 * 4)
            print(f"crappify {dt} {label}: {t*1e3:.0f} us  {nb/t/1e6:.0f} GB/s")
    t64 = torch.randint(0, 255, (64, 512, 512), device="cuda").to(torch.int16)
    tab = ops.TileTable([t64], [0] * 64, list(range(64)), [0] * 64, [0] * 64, [512] * 64, [512] * 64)
    t = ev(lambda: ops.crappify(tab, 512, 4, specs, clip_between=True, want_hr_u8=True))
    print(f"crappify config-2 step (64 x 512^2 u16 + HR u8): {t*1e3:.1f} us  {64*(512*512*3+128*128*4)/t/1e6:.0f} GB/s")
else:
    for env in ({}, {"PSSR_STITCH_NOBAND": "1"}, {"PSSR_CRAP_MINB": "4"}):
        print("==== env", env, flush=True)
        subprocess.run([sys.executable, __file__, "run"], env=dict(os.environ, **env))
