"""Tile-wise sharding across the GPUs of one box.

Tiles are independent (a dataset item is pure given its noise draw; the forward is per-sample in
eval mode; metrics are per-image), so the path shards with NO data-path collective: every rank takes a
contiguous block of the validation items.  ``torch.distributed`` (NCCL over NVLink on GPUs, gloo in the
CPU tests) is used only to (1) all-gather per-image metric values / all-reduce their sums and
(2) gather predicted tiles or stitched sheets on rank 0.  Without an initialised process group every
function degenerates to the single-process case.
"""
import os

import torch
import torch.distributed as dist


def is_dist() -> bool:
    return dist.is_available() and dist.is_initialized()


def rank() -> int:
    return dist.get_rank() if is_dist() else 0


def world_size() -> int:
    return dist.get_world_size() if is_dist() else 1


def init_from_env(backend: str = None):
    """One process per GPU, launched by torchrun: reads RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*."""
    if is_dist() or int(os.environ.get("WORLD_SIZE", "1")) <= 1:
        return rank(), world_size()
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dist.init_process_group(backend=backend)
    return rank(), world_size()


def shard_bounds(n_items: int, r: int, w: int):
    """Contiguous, balanced block [lo, hi) of rank r among w ranks (first n % w ranks get one extra)."""
    base, extra = divmod(n_items, w)
    lo = r * base + min(r, extra)
    return lo, lo + base + (1 if r < extra else 0)


def shard_range(n_items: int):
    return shard_bounds(n_items, rank(), world_size())


def shard_groups(counts, r: int, w: int):
    """Group-aligned sharding ("aligned to whole sheets where possible so each rank stitches its own sheets").
    ``counts[g]`` items belong to group g (a sheet), groups are contiguous in item order.  Returns (first group, end group,
    first item, end item) of rank r: contiguous runs of whole groups, balanced by item count with a greedy prefix split."""
    total = sum(counts)
    bounds, acc, g = [0], 0, 0
    for k in range(1, w):
        target = total * k / w
        while g < len(counts) and acc + counts[g] / 2 <= target:
            acc += counts[g]
            g += 1
        bounds.append(g)
    bounds.append(len(counts))
    g0, g1 = bounds[r], bounds[r + 1]
    return g0, g1, sum(counts[:g0]), sum(counts[:g1])


def allreduce_vector(t: torch.Tensor) -> torch.Tensor:
    """In-place SUM of a small tensor over ranks on its own device (NCCL) or via the CPU (gloo)."""
    if not is_dist():
        return t
    if dist.get_backend() == "nccl":
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t
    c = t.cpu()
    dist.all_reduce(c, op=dist.ReduceOp.SUM)
    t.copy_(c)
    return t


class SheetGather:
    """Stitched uint8 sheets to the HOST of rank 0 (the path's second collective), streamed in rounds: round j moves
    the j-th sheet of every rank.  Every rank calls :meth:`post` right after its j-th sheet has been stitched on the current
    stream; the NCCL point-to-point kernels (NVLink) then run beside the kernels of the next sheet, and on rank 0 a side stream
    waits for each round and copies it to pinned host memory -- one GPU's PCIe link carries every sheet of the job, so the copies
    have to overlap the forward passes instead of queuing up behind the last one (8 x 4 sheets: 10 ms of copies at the end).
    gloo (CPU tests): the same rounds with host tensors."""

    def __init__(self, owners, shapes, device):
        self.owners, self.shapes, self.device = list(owners), list(shapes), device
        self.by_rank = {}
        for s_, o in enumerate(self.owners):
            self.by_rank.setdefault(o, []).append(s_)
        self.n_rounds = max((len(v) for v in self.by_rank.values()), default=0)
        self.active = is_dist() and world_size() > 1
        self.cpu = self.active and dist.get_backend() != "nccl"
        self.side = torch.cuda.Stream(device=device) if (self.active and not self.cpu) else None
        self.hosts = {}            # rank 0: sheet index -> host tensor (pinned under NCCL)
        self._pending, self._keep = [], []

    def post(self, j, local):
        """Round j: ``local`` maps this rank's sheet indices to their (device) tensors; ranks without a j-th sheet just return."""
        if not self.active:
            return
        r, ops_, got = rank(), [], []
        if r == 0:
            for q, sheets in self.by_rank.items():
                if q != 0 and j < len(sheets):
                    buf = torch.empty(tuple(self.shapes[sheets[j]]), dtype=torch.uint8, device="cpu" if self.cpu else self.device)
                    ops_.append(dist.P2POp(dist.irecv, buf, q))
                    got.append((sheets[j], buf))
        else:
            sheets = self.by_rank.get(r, [])
            if j < len(sheets):
                t = local[sheets[j]].contiguous()
                t = t.cpu() if self.cpu else t
                self._keep.append(t)
                ops_.append(dist.P2POp(dist.isend, t, 0))
        if not ops_:
            return
        reqs = dist.batch_isend_irecv(ops_)
        if r != 0:
            self._pending += reqs
        elif self.cpu:
            for req in reqs:
                req.wait()
            for s_, buf in got:
                self.hosts[s_] = buf
        else:
            with torch.cuda.device(self.device), torch.cuda.stream(self.side):
                for req in reqs:
                    req.wait()             # orders the SIDE stream after the transfer; the compute stream never waits for it
                for s_, buf in got:
                    h = torch.empty(buf.shape, dtype=torch.uint8, pin_memory=True)
                    h.copy_(buf, non_blocking=True)
                    buf.record_stream(self.side)
                    self.hosts[s_] = h

    def finish(self):
        """Waits for everything posted; rank 0 gets {sheet index: host tensor} of the other ranks' sheets."""
        for req in self._pending:
            req.wait()
        if self.side is not None:
            self.side.synchronize()
        self._keep.clear()
        return self.hosts


def gather_sheets(local, owners, shapes, device):
    """All rounds of :class:`SheetGather` at once.  ``local``: {sheet index: tensor} of this rank; ``owners[s]``: the rank that
    holds sheet s; ``shapes[s]``: its shape.  Rank 0 returns the full list in sheet order (its own sheets as given, the others on
    the host), the other ranks their own sheets (None elsewhere)."""
    n = len(owners)
    out = [local.get(s) for s in range(n)]
    g = SheetGather(owners, shapes, device)
    for j in range(g.n_rounds):
        g.post(j, local)
    for s_, h in g.finish().items():
        out[s_] = h
    return out


def _device():
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def allreduce_sums(values):
    """Sum a small float64 vector over ranks (metric sums, counts)."""
    t = torch.as_tensor(values, dtype=torch.float64)
    if not is_dist():
        return t
    t = t.to(_device())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu()


def gather_metric_lists(per_image: dict, names):
    """All-gather the per-image metric lists in validation order (rank blocks are contiguous)."""
    if not is_dist():
        return per_image
    gathered = [None] * world_size()
    dist.all_gather_object(gathered, {m: per_image[m] for m in names})
    return {m: [v for g in gathered for v in g[m]] for m in names}


def gather_dict(outs: dict):
    """Predicted tiles name -> uint8 array: rank 0 receives the union (others get their own share)."""
    if not is_dist():
        return outs
    gathered = [None] * world_size() if rank() == 0 else None
    dist.gather_object(outs, gathered, dst=0)
    if rank() == 0:
        merged = {}
        for g in gathered:
            merged.update(g)
        return merged
    return outs


def gather_images_nccl(local: torch.Tensor, n_items: int, counts=None):
    """NCCL path of predict_images: every rank holds the uint8 predictions of its contiguous block of the validation items as
    ONE device tensor [n_local, ...]; rank 0 receives all blocks with one `gather` over NVLink and returns
    [n_items, ...] in validation order (other ranks return None).  Blocks are padded to the largest share.
    ``counts``: items per rank when the shares are not the balanced ``shard_bounds`` split (sheet-aligned sharding)."""
    w, r = world_size(), rank()
    if counts is None:
        counts = [shard_bounds(n_items, k, w)[1] - shard_bounds(n_items, k, w)[0] for k in range(w)]
    mx = max(counts)
    if min(counts) == 0:        # fewer items than ranks: an empty rank does not know the image shape
        shapes = [None] * w
        dist.all_gather_object(shapes, tuple(local.shape[1:]) if local.shape[0] else None)
        shape = next(sh for sh in shapes if sh is not None)
        if local.shape[0] == 0:
            local = torch.zeros((0,) + tuple(shape), dtype=local.dtype, device=local.device)
    buf = local
    if local.shape[0] < mx:
        buf = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        buf[:local.shape[0]] = local
    bufs = [torch.empty_like(buf) for _ in range(w)] if r == 0 else None
    dist.gather(buf.contiguous(), bufs, dst=0)
    if r != 0:
        return None
    return torch.cat([b[:c] for b, c in zip(bufs, counts)], 0)


def gather_tensor_to_rank0(t: torch.Tensor):
    """Gather equally-shaped device tensors (e.g. stitched uint8 sheets) on rank 0 with one collective."""
    if not is_dist():
        return [t]
    if dist.get_backend() == "nccl":
        bufs = [torch.empty_like(t) for _ in range(world_size())]
        dist.all_gather(bufs, t.contiguous())
        return bufs if rank() == 0 else None
    bufs = [torch.empty_like(t) for _ in range(world_size())] if rank() == 0 else None
    dist.gather(t.contiguous(), bufs, dst=0)
    return bufs
