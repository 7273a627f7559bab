"""Developer probe: the bench's e2e call repeated, with the SM clock / power / temperature sampled after each call."""
import sys, time, contextlib, io
import torch
sys.path.insert(0, ".")
import pynvml
from bench import _synthetic_tiles, BATCH, TILE, SCALE
from pssr2_b200.crappifiers import AdditiveGaussian, MultiCrappifier, Poisson
from pssr2_b200.data import ImageDataset
from pssr2_b200.models import ResUNet
from pssr2_b200.predict import predict_images

pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = ResUNet().eval().to(dev)
crap = MultiCrappifier(Poisson(), AdditiveGaussian())
host = [_synthetic_tiles(BATCH, s, dev).cpu().pin_memory() for s in (1, 2)]
steps = 50
stacks = [host[i % 2] for i in range(steps)]
def run():
    ds = ImageDataset(list(stacks), hr_res=TILE, lr_scale=SCALE, crappifier=crap, n_frames=1, val_split=1, device=dev)
    ds.rank_local = True
    return predict_images(model, ds, device=str(dev), batch_size=BATCH, out_dir=None)
keep = len(sys.argv) > 1
with contextlib.redirect_stderr(io.StringIO()):
    p = run(); del p
    for i in range(8):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        p = run()
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        clk = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM); pw = pynvml.nvmlDeviceGetPowerUsage(h) / 1e3
        tmp = pynvml.nvmlDeviceGetTemperature(h, pynvml.NVML_TEMPERATURE_GPU)
        print(f"call {i}: {1e3*dt/steps:.3f} ms/step  sm {clk} MHz  {pw:.0f} W  {tmp} C", flush=True)
        del p
        if i == 3:
            time.sleep(2.0)          # let the board cool: does the next call recover?
