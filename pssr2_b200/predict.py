"""``predict_images`` / ``test_metrics`` with the reference's signatures (pssr/predict.py:11-83, :144-211).

Where the reference walks a DataLoader item by item on the host, these entry points generate each batch
on the device with one fused crappify launch, run the network plan, apply ``_pred_array``
(predict.py:245-246) in the plan's epilogue, normalise and score on the device, and move only uint8
results / a handful of sums back to the host.  Under ``torchrun`` (see pssr2_b200/dist.py) the
validation items are sharded tile-wise across ranks.
"""
import math
import os
from collections.abc import Mapping

import numpy as np
import torch
import torch.nn as nn

from . import dist as D
from . import ops
from .data import _DeviceDataset
from .models import _PlanModule
from .util import _get_callbacks, pixel_metric


def _progress(it, **kw):
    try:
        from tqdm import tqdm
        return tqdm(it, **kw)
    except Exception:
        return it


def _pred_u8(model, lr):
    """model(lr) followed by `_pred_array` (clip -> uint8 truncation -> centre channel), on the device."""
    if isinstance(model, _PlanModule):
        _, out8 = model.forward_u8(lr)
        return out8
    out = model(lr)  # a foreign nn.Module: its own forward, then the same epilogue
    c = out.shape[1] // 2
    return out[:, c:c + 1].clamp(0, 255).to(torch.uint8)


def _batch(dataset, idxs, device, want_hr_u8, tile_index0=None):
    """-> (lr float32 [n,C,h,w] on device, hr uint8 [n,1,H,W] on device or None)."""
    if isinstance(dataset, _DeviceDataset):
        b = dataset.batch(idxs, want_hr=False, want_hr_u8=want_hr_u8, tile_index0=tile_index0)
        return b["lr"], b["hr_u8"]
    # duck-typed foreign dataset (reference contract: __getitem__ -> (hr, lr) or lr CPU tensors)
    items = [dataset[i] for i in idxs]
    if dataset.is_lr:
        return torch.stack([torch.as_tensor(i) for i in items]).to(device), None
    hr = torch.stack([torch.as_tensor(i[0]) for i in items]).to(device)
    lr = torch.stack([torch.as_tensor(i[1]) for i in items]).to(device)
    c = hr.shape[1] // 2
    return lr, (hr[:, c:c + 1].clamp(0, 255).to(torch.uint8) if want_hr_u8 else None)


def _shares(dataset, n_items):
    """Item range [lo, hi) of this rank and the item count of every rank.  Sliding datasets are split at sheet boundaries when
    there are at least as many sheets as ranks, so that a rank only uploads, predicts and stitches its own sheets (SURVEY.md
    8e); otherwise contiguous balanced blocks of tiles."""
    w, r = D.world_size(), D.rank()
    runs = dataset.sheet_item_counts() if (w > 1 and hasattr(dataset, "sheet_item_counts") and hasattr(dataset, "tiles")) else None
    if runs is not None and len(runs) >= w and sum(c for _, c in runs) == n_items:
        counts = [c for _, c in runs]
        spans = [D.shard_groups(counts, k, w) for k in range(w)]
        return spans[r][2], spans[r][3], [sp[3] - sp[2] for sp in spans]
    spans = [D.shard_bounds(n_items, k, w) for k in range(w)]
    return spans[r][0], spans[r][1], [hi - lo for lo, hi in spans]


class TilePreds(Mapping):
    """What ``predict_images(..., keep_on_device=True)`` returns: the reference's ``dict[name -> uint8 [1, H, W]]`` view of
    predictions that stay in HBM.  ``reassemble_sheets`` reads the device batches directly (no tile ever visits the host);
    indexing by name copies that batch to the host on first use."""

    def __init__(self):
        self.names = []            # validation order
        self.batches = []          # device uint8 [n, 1, H, W]
        self._where = {}           # name -> (batch index, row)
        self._host = {}

    def append(self, names, batch):
        for k, nme in enumerate(names):
            self._where[nme] = (len(self.batches), k)
        self.names += list(names)
        self.batches.append(batch)

    def device_tiles(self, names):
        """uint8 [len(names), H, W] on the device, in the given order (one gather, or a view when the names are one batch run)."""
        loc = [self._where[nme] for nme in names]
        b0 = loc[0][0]
        if all(b == b0 for b, _ in loc) and [k for _, k in loc] == list(range(loc[0][1], loc[0][1] + len(loc))):
            return self.batches[b0][loc[0][1]:loc[0][1] + len(loc), 0]
        return torch.stack([self.batches[b][k, 0] for b, k in loc])

    def __getitem__(self, name):
        b, k = self._where[name]
        if b not in self._host:
            self._host[b] = self.batches[b].cpu().numpy()
        return self._host[b][k]

    def __iter__(self):
        return iter(self.names)

    def __len__(self):
        return len(self.names)


def _to_device(model, device):
    """model.to(device) without walking every parameter when the model already lives there."""
    want, p = torch.device(device), next(model.parameters(), None)
    same = p is not None and p.device.type == want.type and (want.index is None or p.device.index == want.index)
    if not same:
        model.to(device)


def predict_images(model: nn.Module, dataset, device: str = "cuda", batch_size=None, out_dir: str = "preds", norm: bool = False,
                   prefix: str = None, dataloader_kwargs=None, callbacks=None, keep_on_device: bool = False):
    r"""Predicts high-resolution images from low-resolution images (pssr/predict.py:11-83).

    Same arguments and return value as the reference (``dict[name -> uint8 [1,H,W]]`` iff ``out_dir`` is None,
    else ``{out_dir}/{prefix_}{name}.tif`` files).  ``device`` must be a CUDA device; ``dataloader_kwargs`` is
    accepted for compatibility (there is no DataLoader: batches are generated on the device).
    Under torchrun each rank predicts a contiguous share of ``dataset.val_idx`` (whole sheets of a sliding dataset where
    possible) and rank 0 receives all images.  ``keep_on_device=True`` (with ``out_dir=None``) returns a :class:`TilePreds`
    mapping whose tiles stay in HBM -- every rank its own share -- for :func:`pssr2_b200.util.reassemble_sheets`."""
    batch_size = 1 if batch_size is None else batch_size
    if norm and dataset.is_lr:
        raise ValueError("Dataset must be paired with high-low-resolution images for normalization.")
    if not str(device).startswith("cuda"):
        raise RuntimeError("pssr2_b200.predict_images runs on CUDA devices only (no CPU fallback); pass device='cuda'")
    writer = None
    if out_dir:
        os.makedirs(out_dir, exist_ok=True)
        from .io import TiffWriter
        writer = TiffWriter()
    callbacks, callback_locals = _get_callbacks(callbacks)
    _to_device(model, device)
    model.eval()

    val_idx = list(dataset.val_idx)
    # `dataset.rank_local = True`: the dataset already holds only this rank's share (pre-sharded ingest, weak scaling): no
    # sharding of val_idx and no gather -- every rank returns / writes its own images
    rank_local = bool(getattr(dataset, "rank_local", False))
    lo, hi, counts = (0, len(val_idx), [len(val_idx)]) if rank_local else _shares(dataset, len(val_idx))
    outs = {}
    if keep_on_device and (out_dir or callbacks):
        raise ValueError("keep_on_device=True returns device-resident tiles: it needs out_dir=None and no callbacks")
    dpreds = TilePreds() if keep_on_device else None
    dev = torch.device(device)
    cur = torch.cuda.current_stream(dev)
    down = torch.cuda.Stream(device=dev)          # device -> pinned host copies run beside the next batch's kernels
    pending = None                                # (done event, pinned uint8 batch, positions, device buffer)

    def finish(item):
        done, host, pos, _keep = item
        done.synchronize()
        hr_hat = host.numpy()                     # zero-copy view of the pinned batch; the dict entries keep it alive
        for batch_idx, image_idx in enumerate(pos):
            name = dataset._get_name(image_idx)   # reference quirk: the POSITION in val_idx names the file (predict.py:69-73)
            if out_dir:
                writer.write(f"{out_dir}/{prefix + '_' if prefix else ''}{name}.tif", hr_hat[batch_idx])   # encoded off-thread
            else:
                outs[name] = hr_hat[batch_idx]
            for idx, callback in enumerate(callbacks):
                if callback_locals[idx]:
                    callback(locals())
                else:
                    callback()

    # under torchrun with NCCL the predicted tiles travel GPU -> GPU: every rank keeps its uint8 batches on the device, rank 0
    # gathers them with one collective over NVLink and reads them back once (no pickling of host arrays)
    nccl_gather = (out_dir is None and not rank_local and not keep_on_device and D.is_dist() and torch.distributed.get_backend() == "nccl"
                   and D.world_size() > 1 and not callbacks)
    kept = []
    starts = range(lo, hi, batch_size)
    with torch.no_grad():
        for start in (_progress(starts) if len(starts) > 1 else starts):
            pos = list(range(start, min(start + batch_size, hi)))
            idxs = [val_idx[p] for p in pos]
            lr, hr8 = _batch(dataset, idxs, device, want_hr_u8=norm)
            hr_hat = _pred_u8(model, lr)
            if norm:
                _, hr_hat = ops.normalize_preds_u8(hr8[:, 0], hr_hat[:, 0])
                hr_hat = hr_hat[:, None]
            crop_res = dataset.crop_res if not dataset.is_lr else dataset.crop_res * (hr_hat.shape[-1] // lr.shape[-1])
            # own copy of the batch: the plan's output buffer is overwritten by the next forward while this one travels
            dbuf = hr_hat[:, :, :crop_res, :crop_res].clone(memory_format=torch.contiguous_format)
            if keep_on_device:
                dpreds.append([dataset._get_name(p) for p in pos], dbuf)
                continue
            if nccl_gather:
                kept.append(dbuf)
                continue
            host = torch.empty(dbuf.shape, dtype=torch.uint8, pin_memory=True)
            ready = torch.cuda.Event()
            ready.record(cur)
            done = torch.cuda.Event()
            with torch.cuda.stream(down):
                down.wait_event(ready)
                host.copy_(dbuf, non_blocking=True)
                done.record(down)
            if pending is not None:
                finish(pending)                   # the previous batch reaches the host while this one computes
            pending = (done, host, pos, dbuf)
        if pending is not None:
            finish(pending)
    if writer is not None:
        writer.close()                    # every file is on disk when the call returns
    if keep_on_device:
        return dpreds
    if nccl_gather:
        shape = kept[0].shape[1:] if kept else (1, dataset.crop_res, dataset.crop_res)
        local = torch.cat(kept, 0) if kept else torch.zeros((0,) + tuple(shape), dtype=torch.uint8, device=dev)
        allp = D.gather_images_nccl(local, len(val_idx), counts)
        src, base = (allp, 0) if allp is not None else (local, lo)          # rank 0: everything; others: their own share
        arr = src.cpu().numpy()
        return {dataset._get_name(base + k): arr[k] for k in range(arr.shape[0])}
    if out_dir is None:
        return outs if rank_local else D.gather_dict(outs)


def _collage_preds(lr8, hr_hat8, hr8, norm: bool = False, max_images: int = 5, crop_res: int = None, lr_scale: int = 4):
    """pssr/predict.py:213-243 on uint8 device tensors [n,1,h,w] (what `_pred_array` yields): crop, normalise on the device
    (hr_hat against hr, then the low-resolution input against the normalised hr -- the differing-resolution path of
    `normalize_preds`, util.py:179), then the reference's Pillow layout: images stacked vertically, the three columns
    (input enlarged with NEAREST | prediction | ground truth) side by side."""
    from PIL import Image
    from .util import normalize_preds
    crop_res = hr_hat8.shape[-1] if crop_res is None else crop_res
    lr_scale = int(hr_hat8.shape[-1] / lr8.shape[-1]) if lr_scale is None else lr_scale
    lr8 = lr8[:, :, :crop_res // lr_scale, :crop_res // lr_scale].contiguous()
    hr_hat8 = hr_hat8[:, :, :crop_res, :crop_res].contiguous()
    hr8 = None if hr8 is None else hr8[:, :, :crop_res, :crop_res].contiguous()
    if norm:
        hr8, hr_hat8 = normalize_preds(hr8, hr_hat8)
        _, lr8 = normalize_preds(hr8, lr8)

    def stack(data):
        images = [Image.fromarray(im) for im in data[:min(max_images, len(data)), 0].cpu().numpy()]
        width, height = images[0].width, images[0].height
        out = Image.new("L", (width, height * len(images)))
        for k, im in enumerate(images):
            out.paste(im, (0, height * k))
        return out

    lr_im, hat_im, hr_im = stack(lr8), stack(hr_hat8), None if hr8 is None else stack(hr8)
    lr_im = lr_im.resize((hat_im.width, hat_im.height), Image.Resampling.NEAREST)
    if hr_im is not None and hat_im.size != hr_im.size:
        hat_im = hat_im.resize((hr_im.width, hr_im.height), Image.Resampling.NEAREST)
    cols = [lr_im, hat_im] + ([hr_im] if hr_im is not None else [])
    row = Image.new("L", (cols[0].width * len(cols), cols[0].height))
    for k, im in enumerate(cols):
        row.paste(im, (cols[0].width * k, 0))
    return row


def predict_collage(model: nn.Module, dataset, device: str = "cuda", norm: bool = True, n_images: int = None, prefix: str = None,
                    out_dir: str = "preds", callbacks=None):
    r"""Saves an image collage of vertically stacked rows of the low-resolution input, the PSSR prediction and the high-resolution
    ground truth, in that order (pssr/predict.py:85-142).  Crappification, forward pass, `_pred_array` and the normalisation run on
    the device; only the uint8 rows travel to the host, where Pillow lays them out and writes the PNG like the reference.
    Only evaluation images are used; the order is the reference's (``np.random.seed(0)`` shuffle when ``val_split < 1``)."""
    from PIL import Image
    if norm and dataset.is_lr:
        raise ValueError("Dataset must be paired with high-low-resolution images for normalization.")
    if not str(device).startswith("cuda"):
        raise RuntimeError("pssr2_b200.predict_collage runs on CUDA devices only (no CPU fallback); pass device='cuda'")
    callbacks, callback_locals = _get_callbacks(callbacks)
    n_images = min(50, len(dataset)) if n_images is None else n_images
    _to_device(model, device)
    model.eval()

    order = list(dataset.val_idx)
    if len(dataset.val_idx) < len(dataset):          # only shuffle if val_split < 1 (data.py:737-749, seed=True)
        np.random.seed(0)
        np.random.shuffle(order)
    collage = Image.new("L", (dataset.crop_res * (2 if dataset.is_lr else 3), dataset.crop_res * n_images))
    with torch.no_grad():
        for idx, data_idx in enumerate(order):
            lr, hr8 = _batch(dataset, [data_idx], device, want_hr_u8=not dataset.is_lr, tile_index0=idx)
            hr_hat8 = _pred_u8(model, lr)
            c = lr.shape[1] // 2
            lr8 = lr[:, c:c + 1].clamp(0, 255).to(torch.uint8)         # `_pred_array` of the input
            collage.paste(_collage_preds(lr8, hr_hat8, hr8, norm, 1, dataset.crop_res, dataset.lr_scale), (0, dataset.crop_res * idx))
            for idx, callback in enumerate(callbacks):                 # (sic: the reference reuses `idx` here, predict.py:132)
                if callback_locals[idx]:
                    callback(locals())
                else:
                    callback()
            if idx >= n_images - 1:
                break
    os.makedirs(out_dir, exist_ok=True)
    collage.save(f"{out_dir}/{prefix + '_' if prefix else ''}collage_{n_images}.png")



def test_metrics(model: nn.Module, dataset, device: str = "cuda", metrics=["mse", "pixel", "psnr", "ssim"], avg: bool = True,
                 norm: bool = True, callbacks=None, batch_size: int = 1, item0_quirk: bool = True):
    r"""Computes restoration metrics of predicted vs ground truth images (pssr/predict.py:144-211).

    ``item0_quirk=True`` (default) reproduces the reference, which evaluates ``dataset[0]`` -- re-crappified with
    fresh noise -- once per validation index (predict.py:180); ``False`` scores every validation item.
    ``batch_size`` only groups launches; results do not depend on it.  Under torchrun the items are sharded and
    the per-image values are all-gathered (``avg=False``) / their sums all-reduced (``avg=True``)."""
    callbacks, callback_locals = _get_callbacks(callbacks)
    image_range = 255
    metrics = [metrics] if type(metrics) is str else metrics
    names = list(metrics)
    if not str(device).startswith("cuda"):
        raise RuntimeError("pssr2_b200.test_metrics runs on CUDA devices only (no CPU fallback); pass device='cuda'")
    _to_device(model, device)
    model.eval()

    val_idx = list(dataset.val_idx)
    lo, hi = D.shard_range(len(val_idx))      # item-0 quirk / per-tile metrics: sheets play no role, balanced blocks
    per_image = {m: [] for m in names}
    want_ssim = "ssim" in names
    dev = torch.device(device)
    cur = torch.cuda.current_stream(dev)
    pending = None                                # (done event, pinned sums [2, n] float64, n, pixels, ssim pixels)

    def finish(item):
        done, host, n, n_px, n_win = item
        done.synchronize()
        sq, ss = host[0].numpy(), host[1].numpy()
        for i in range(n):
            # mean((h/255 - h_hat/255)^2) in float64 == sum d^2 / N / 255^2 up to 1e-16 relative
            mse = float(sq[i]) / n_px / float(image_range) ** 2
            if "mse" in per_image:
                per_image["mse"].append(mse)
            if "pixel" in per_image:
                per_image["pixel"].append(pixel_metric(mse, image_range))
            if "psnr" in per_image:
                err = float(sq[i]) / n_px
                per_image["psnr"].append(10 * math.log10(image_range ** 2 / err) if err > 0 else float("inf"))
            if "ssim" in per_image:
                per_image["ssim"].append(float(ss[i]) / n_win)
        for idx, callback in enumerate(callbacks):
            if callback_locals[idx]:
                callback(locals())
            else:
                callback()

    starts = range(lo, hi, batch_size)
    with torch.no_grad():
        for start in (_progress(starts) if len(starts) > 1 else starts):
            pos = list(range(start, min(start + batch_size, hi)))
            idxs = [0] * len(pos) if item0_quirk else [val_idx[p] for p in pos]
            lr, hr8 = _batch(dataset, idxs, device, want_hr_u8=True, tile_index0=pos[0])
            hr_hat = _pred_u8(model, lr)
            crop_res = dataset.crop_res if not dataset.is_lr else dataset.crop_res * (hr_hat.shape[-1] // lr.shape[-1])
            hr, hr_hat = hr8[:, 0, :crop_res, :crop_res].contiguous(), hr_hat[:, 0, :crop_res, :crop_res].contiguous()
            if norm:
                hr, hr_hat = ops.normalize_preds_u8(hr, hr_hat)
            sq, ss = ops.metric_sums(hr, hr_hat, want_ssim=want_ssim)
            # the two small sum vectors travel to pinned memory asynchronously; they are read one batch later
            both = torch.stack([sq.to(torch.float64), (ss if ss is not None else torch.zeros_like(sq)).to(torch.float64)])
            host = torch.empty(both.shape, dtype=torch.float64, pin_memory=True)
            host.copy_(both, non_blocking=True)
            done = torch.cuda.Event()
            done.record(cur)
            if pending is not None:
                finish(pending)
            pending = (done, host, len(pos), hr.shape[-1] * hr.shape[-2], (hr.shape[-2] - 6) * (hr.shape[-1] - 6))
        if pending is not None:
            finish(pending)
    if avg:
        # the path's first collective: one all-reduce of [sum per metric ..., count] in float64
        vec = torch.tensor([sum(per_image[m]) for m in names] + [float(hi - lo)], dtype=torch.float64, device=dev)
        vec = D.allreduce_vector(vec).cpu()
        return {m: float(vec[i]) / float(vec[-1]) for i, m in enumerate(names)}
    per_image = D.gather_metric_lists(per_image, names)
    return {m: v for m, v in per_image.items()}


test_metrics.__test__ = False  # "This guy is NOT a test." (reference tests/conftest.py:1-2)


def predict_sheets(model: nn.Module, dataset, device: str = "cuda", batch_size: int = None, overlap: int = None, margin: int = 0,
                   out_dir: str = None, norm: bool = False, prefix: str = None):
    r"""``predict_images`` + ``reassemble_sheets`` (pssr/predict.py:11-83, pssr/util.py:54-108) as ONE device-resident pipeline
    for sliding datasets: every rank uploads, predicts and stitches its own sheets -- the tiles never leave HBM -- and only the
    stitched uint8 sheets travel (device -> pinned host on a side stream; under torchrun rank 0 receives every sheet over NCCL).

    ``overlap`` / ``margin``: as in ``reassemble_sheets`` with ``lr_scale=1`` on the predicted tiles, in pixels of the
    PREDICTION (default overlap: the dataset's own tile overlap, scaled in LR mode).  Returns the list of stitched sheets
    ``uint8 [stacks, rows*step+overlap, cols*step+overlap]`` in sheet order (rank 0: all sheets; other ranks: a list of None,
    their sheets travel to rank 0) or writes ``{out_dir}/{prefix_}{sheet}.tif`` (rank 0)."""
    if not hasattr(dataset, "tiles") or not hasattr(dataset, "_tiles_y"):
        raise TypeError("predict_sheets needs a SlidingDataset (tiled sheets)")
    if not str(device).startswith("cuda"):
        raise RuntimeError("pssr2_b200.predict_sheets runs on CUDA devices only (no CPU fallback); pass device='cuda'")
    if norm and dataset.is_lr:
        raise ValueError("Dataset must be paired with high-low-resolution images for normalization.")
    runs = dataset.sheet_item_counts()
    n_val = len(dataset.val_idx)
    if runs is None or any(c != dataset.tiles[i] * dataset.slices[i] for i, c in runs):
        raise ValueError("predict_sheets stitches whole sheets: construct the dataset with val_split=1")
    batch_size = 1 if batch_size is None else batch_size
    if out_dir:
        os.makedirs(out_dir, exist_ok=True)
    _to_device(model, device)
    model.eval()
    dev = torch.device(device)
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    w, r = D.world_size(), D.rank()
    counts = [c for _, c in runs]
    spans = [D.shard_groups(counts, k, w) for k in range(w)] if len(runs) >= w else None
    if spans is None:
        raise ValueError(f"predict_sheets shards by whole sheets: {len(runs)} sheets cannot feed {w} ranks")
    owners = [next(k for k in range(w) if spans[k][0] <= g < spans[k][1]) for g in range(len(runs))]
    g0, g1, lo, _ = spans[r]
    val_idx = list(dataset.val_idx)
    down = torch.cuda.Stream(device=dev)
    local, shapes, hosts = {}, [None] * len(runs), {}
    # every rank knows every sheet's stitched shape (rank 0 posts the receives of a round before the senders are done)
    for g in range(len(runs)):
        img = runs[g][0]
        ty = dataset._tiles_y[img]
        tx = dataset.tiles[img] // ty
        T = dataset.crop_res * (1 if not dataset.is_lr else max(1, getattr(model, "scale", 1)))
        ovg = (dataset.hr_res - dataset.stride) * (T // dataset.crop_res) if overlap is None else overlap
        shapes[g] = (dataset.slices[img], tx * (T - ovg) + ovg, ty * (T - ovg) + ovg)
    gather = D.SheetGather(owners, shapes, dev)
    pos = lo
    with torch.no_grad(), torch.cuda.device(dev):
        cur = torch.cuda.current_stream(dev)
        for g in range(len(runs)):
            img, cnt = runs[g]
            n_sl = dataset.slices[img]
            ty = dataset._tiles_y[img]
            tx = dataset.tiles[img] // ty
            scale_out = 1
            if g0 <= g < g1:
                tiles_dev = None
                for start in range(pos, pos + cnt, batch_size):
                    ps = list(range(start, min(start + batch_size, pos + cnt)))
                    lr, hr8 = _batch(dataset, [val_idx[p] for p in ps], device, want_hr_u8=norm)
                    hr_hat = _pred_u8(model, lr)
                    if norm:
                        _, hh = ops.normalize_preds_u8(hr8[:, 0], hr_hat[:, 0])
                        hr_hat = hh[:, None]
                    crop_res = dataset.crop_res if not dataset.is_lr else dataset.crop_res * (hr_hat.shape[-1] // lr.shape[-1])
                    if tiles_dev is None:
                        tiles_dev = torch.empty(cnt, crop_res, crop_res, dtype=torch.uint8, device=dev)
                        scale_out = hr_hat.shape[-1] // lr.shape[-1] if dataset.is_lr else 1
                    tiles_dev[start - pos:start - pos + len(ps)] = hr_hat[:, 0, :crop_res, :crop_res]
                pos += cnt
                # item order inside a sheet is tile-major, slice-minor (data.py:236-256); the stitch wants [slice][tile]
                if n_sl > 1:
                    tiles_dev = tiles_dev.view(tx * ty, n_sl, *tiles_dev.shape[1:]).transpose(0, 1).reshape(cnt, *tiles_dev.shape[1:])
                ov = (dataset.hr_res - dataset.stride) * scale_out if overlap is None else overlap
                sheet = ops.stitch(tiles_dev, tx, ty, ov, margin)
                local[g] = sheet
                if r == 0 or not D.is_dist():
                    # rank 0's own sheets go to pinned host memory beside the next sheet's kernels
                    host = torch.empty(sheet.shape, dtype=torch.uint8, pin_memory=True)
                    ready = torch.cuda.Event()
                    ready.record(cur)
                    with torch.cuda.stream(down):
                        down.wait_event(ready)
                        host.copy_(sheet, non_blocking=True)
                    sheet.record_stream(down)
                    hosts[g] = host
                # round j of the sheet gather: this rank's j-th sheet travels to rank 0 (NVLink -> rank 0's pinned host memory)
                # beside the kernels of its next sheet (a pageable `.cpu()` per gathered sheet after the last forward cost rank 0
                # 90 ms at 8 GPUs x 2 sheets, for 41 ms of forward time)
                gather.post(g - g0, local)
        for j in range(g1 - g0, gather.n_rounds):      # ranks with fewer sheets: the later rounds only concern rank 0's receives
            gather.post(j, local)
    hosts.update(gather.finish())
    down.synchronize()
    out = []
    for g in range(len(runs)):
        if g in hosts:
            out.append(hosts[g].numpy())
        else:                                       # ranks > 0: their sheets have left for rank 0 (a pageable read-back of a
                                                    # 3968^2 sheet costs 7 ms -- a third of its forward time -- for a copy nobody uses)
            out.append(None)
    if out_dir:
        from .io import write_tiff
        for g, arr in enumerate(out):
            if arr is not None:
                name = dataset.hr_files[runs[g][0]].split("/")[-1].split(".")[0]
                write_tiff(f"{out_dir}/{prefix + '_' if prefix else ''}{name}.tif", arr)
        return None
    return out
