"""Multi-GPU path (needs >= 2 GPUs, skipped otherwise): predict_images / test_metrics under torchrun shard the
validation tiles across ranks, and the gathered result must equal the single-process result exactly (noise is keyed by
the global tile index; crappifier=None here so predictions are deterministic)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, pickle
sys.path.insert(0, %r)
import numpy as np, torch
from pssr2_b200 import dist as D
from pssr2_b200.data import SlidingDataset
from pssr2_b200.models import ResUNet
from pssr2_b200.predict import predict_images, predict_sheets, test_metrics
r, w = D.init_from_env("nccl")
dev = f"cuda:{int(os.environ.get('LOCAL_RANK', 0))}"
torch.manual_seed(3)
model = ResUNet(hidden=[64, 128, 256], depth=1).eval()
rng = np.random.default_rng(0)
sheet = rng.poisson(90, (1, 640, 832)).clip(0, 255).astype(np.uint8)
ds = SlidingDataset({"s0": sheet}, hr_res=256, lr_scale=4, overlap=64, val_split=1, crappifier=None, device=dev)
preds = predict_images(model, ds, device=dev, batch_size=2, out_dir=None)
m = test_metrics(model, ds, device=dev, avg=False, item0_quirk=False, norm=False, batch_size=2)
mavg = test_metrics(model, ds, device=dev, avg=True, item0_quirk=False, norm=False, batch_size=2)      # the [sums, count] all-reduce
# three sheets, sheet-aligned shares: every rank uploads / predicts / stitches its own sheets, rank 0 receives them over NCCL
sheets3 = {f"m{i}": np.random.default_rng(10 + i).poisson(90, (1, 448, 640)).clip(0, 255).astype(np.uint8) for i in range(3)}
ds3 = SlidingDataset(sheets3, hr_res=256, lr_scale=4, overlap=64, val_split=1, crappifier=None, device=dev, preload=False)
st = predict_sheets(model, ds3, device=dev, batch_size=3, margin=8)
resident = [i for i, s in enumerate(ds3._sheets) if s is not None]
p3 = predict_images(model, SlidingDataset(sheets3, hr_res=256, lr_scale=4, overlap=64, val_split=1, crappifier=None, device=dev), device=dev,
                    batch_size=4, out_dir=None)
allres = [None] * w
torch.distributed.all_gather_object(allres, resident) if w > 1 else None
if r == 0:
    with open(sys.argv[1], "wb") as f:
        pickle.dump({"keys": sorted(preds), "sum": {k: int(v.astype(np.int64).sum()) for k, v in preds.items()}, "m": m, "mavg": mavg, "world": w,
                     "sheets": [s.tobytes() for s in st], "sheet_shapes": [s.shape for s in st], "resident": allres if w > 1 else [resident],
                     "p3": {k: int(v.astype(np.int64).sum()) for k, v in p3.items()}}, f)
if w > 1:
    torch.distributed.barrier()
'''


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_predict_equals_single(tmp_path):
    import pickle
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    outs = {}
    for n in (1, 2):
        out = tmp_path / f"out{n}.pkl"
        env = dict(os.environ, MASTER_ADDR="127.0.0.1")
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
               "--master-port", str(29600 + n), str(script), str(out)]
        r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
        outs[n] = pickle.load(open(out, "rb"))
    assert outs[2]["world"] == 2
    assert outs[1]["keys"] == outs[2]["keys"] and len(outs[1]["keys"]) == 12
    assert outs[1]["sum"] == outs[2]["sum"]
    for k in outs[1]["m"]:
        assert outs[1]["m"][k] == outs[2]["m"][k]
        assert abs(outs[1]["mavg"][k] - outs[2]["mavg"][k]) <= 1e-12 * max(1.0, abs(outs[1]["mavg"][k]))
        assert abs(outs[1]["mavg"][k] - sum(outs[1]["m"][k]) / len(outs[1]["m"][k])) <= 1e-12 * max(1.0, abs(outs[1]["mavg"][k]))
    # stitched sheets: identical bytes, gathered on rank 0; each rank kept only its own sheets (+ at most one prefetched) resident
    assert outs[1]["sheet_shapes"] == outs[2]["sheet_shapes"] and len(outs[2]["sheets"]) == 3
    assert outs[1]["sheets"] == outs[2]["sheets"]
    # (three sheets of six tiles on two ranks: rank 0 owns sheets 0-1 and may have prefetched sheet 2, rank 1 touches sheet 2 only)
    assert outs[2]["resident"][1] == [2] and set(outs[2]["resident"][0]) >= {0, 1}
    assert outs[1]["p3"] == outs[2]["p3"] and len(outs[2]["p3"]) == 18
