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
# the path's two collectives in their product form: the [sums..., count] all-reduce of test_metrics(avg=True) and the stitched-
# sheet gather of predict_sheets (three sheets of different shapes, rank 1 owns the last two)
vec = D.allreduce_vector(torch.tensor([1.0 + r, 2.0, float(r)], dtype=torch.float64))
assert vec.tolist() == [3.0, 4.0, 1.0]
shapes = [(1, 4, 5), (2, 3, 3), (1, 2, 7)]
owners = [0, 1, 1]
local = {s: torch.full(shapes[s], 10 * s + 1, dtype=torch.uint8) for s in range(3) if owners[s] == r}
got = D.gather_sheets(local, owners, shapes, torch.device("cpu"))
if r == 0:
    assert all(g.shape == shapes[s] and int(g.flatten()[0]) == 10 * s + 1 for s, g in enumerate(got))
else:
    assert got[0] is None and got[1] is not None and got[2] is not None
g0, g1, lo3, hi3 = D.shard_groups([100, 100, 100], r, w)
assert (g0, g1, lo3, hi3) == ((0, 2, 0, 200) if r == 0 else (2, 3, 200, 300)) or (g0, g1) in ((0, 1), (1, 3))
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


def test_shard_groups_properties():
    """Sheet-aligned sharding: every group lands on exactly one rank, runs are contiguous, and the item load is as even as whole
    groups allow (never worse than one group off the ideal prefix split)."""
    from pssr2_b200.dist import shard_groups
    for counts in ([100] * 8, [100] * 3, [5, 50, 5, 50, 5], [1, 2, 3, 4, 5, 6, 7], [10], [7, 0, 7]):
        for w in (1, 2, 3, 8):
            if w > len(counts):
                continue
            spans = [shard_groups(counts, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == len(counts)
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert all(s[2] == sum(counts[:s[0]]) and s[3] == sum(counts[:s[1]]) for s in spans)
            ideal = sum(counts) / w
            assert all(abs(spans[k][3] - ideal * (k + 1)) <= max(counts) for k in range(w))
    spans = [shard_groups([100] * 8, r, 8) for r in range(8)]
    assert [s[3] - s[2] for s in spans] == [100] * 8
