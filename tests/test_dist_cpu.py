"""world_size-2 gloo test of the multi-rank host logic (sharding bounds, metric gathering, dict gather)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, %r)
import torch, torch.distributed as dist
from pssr2_b200 import dist as D
r, w = D.init_from_env("gloo")
assert w == 2 and D.is_dist()
n = 11
lo, hi = D.shard_range(n)
cover = [None, None]
dist.all_gather_object(cover, (lo, hi))
assert cover[0][0] == 0 and cover[0][1] == cover[1][0] and cover[1][1] == n, cover
per = {"mse": [float(i) for i in range(lo, hi)], "ssim": [i * 0.5 for i in range(lo, hi)]}
full = D.gather_metric_lists(per, ["mse", "ssim"])
assert full["mse"] == [float(i) for i in range(n)] and full["ssim"] == [i * 0.5 for i in range(n)]
s = D.allreduce_sums([float(hi - lo), 1.0])
assert s.tolist() == [float(n), 2.0]
outs = D.gather_dict({f"t{i}": i for i in range(lo, hi)})
if r == 0:
    assert sorted(outs) == sorted(f"t{i}" for i in range(n))
bufs = D.gather_tensor_to_rank0(torch.full((3,), r, dtype=torch.uint8))
if r == 0:
    assert [int(b[0]) for b in bufs] == [0, 1]
# image gather (the NCCL path of predict_images; the logic is backend-agnostic): uneven shares, then an empty rank
for n_items in (11, 1):
    lo2, hi2 = D.shard_range(n_items)
    local = torch.arange(lo2, hi2, dtype=torch.uint8).view(-1, 1, 1, 1).expand(-1, 1, 2, 3).contiguous()
    if hi2 == lo2:
        local = torch.zeros(0, dtype=torch.uint8)
    allp = D.gather_images_nccl(local, n_items)
    if r == 0:
        assert allp.shape == (n_items, 1, 2, 3) and allp[:, 0, 0, 0].tolist() == list(range(n_items)), allp.shape
    else:
        assert allp is None
dist.barrier()
sys.stdout.write("rank" + str(r) + "-ok\n"); sys.stdout.flush()
'''


def test_two_rank_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29531")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29531", str(script)], env=env, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "rank0-ok" in r.stdout and "rank1-ok" in r.stdout


def test_shard_bounds_properties():
    from pssr2_b200.dist import shard_bounds
    for n in (0, 1, 7, 64, 100):
        for w in (1, 2, 3, 8):
            edges = [shard_bounds(n, r, w) for r in range(w)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1
