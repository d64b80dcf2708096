"""Data-parallel sharding of the path across the GPUs of one box.

Images are independent and every ROI reads only its own image (``batch_idx``, hed/dynamic_roi_align.py:80,156), so
the batch is split into contiguous image ranges -- balanced on ROI count, since the head dominates the cost -- and each
rank runs the whole path on its range with ``batch_idx`` rebased.  There is NO collective on the hot path; the only
communication is an optional gather of the results for the caller (``gather_logits``).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch


def partition_images(roi_batch_idx: torch.Tensor, n_images: int, world_size: int, roi_cost: float = 2.0, image_cost: float = 1.0) -> List[Tuple[int, int]]:
    """Contiguous image ranges [lo, hi) per rank minimising the maximum cost, cost = image_cost*images + roi_cost*rois
    (per-ROI head ~58 GFLOP vs ~28 GFLOP per 480x640 image for B0)."""
    counts = torch.bincount(roi_batch_idx.long().clamp(0, max(n_images - 1, 0)), minlength=n_images).tolist() if n_images else []
    cost = [image_cost + roi_cost * c for c in counts]
    total = sum(cost)
    bounds, lo, acc = [], 0, 0.0
    for r in range(world_size):
        target = total * (r + 1) / world_size
        hi = lo
        while hi < n_images and (acc + cost[hi] <= target + 1e-9 or hi == lo and n_images - hi >= world_size - r):
            acc += cost[hi]; hi += 1
        if r == world_size - 1:
            hi = n_images
        bounds.append((lo, hi))
        lo = hi
    return bounds


def shard_batch(images: torch.Tensor, rois: torch.Tensor, world_size: int, rank: int, bounds: Sequence[Tuple[int, int]] = None):
    """Returns (images_shard, rois_shard with batch_idx rebased, roi_index) for ``rank``; ``roi_index`` are the positions
    of the shard's ROIs in the original ``rois`` (to restore the caller's order after a gather)."""
    bounds = bounds or partition_images(rois[:, 0], images.shape[0], world_size)
    lo, hi = bounds[rank]
    b = rois[:, 0].long()
    sel = ((b >= lo) & (b < hi)).nonzero(as_tuple=True)[0]
    r = rois[sel].clone()
    r[:, 0] -= lo
    return images[lo:hi], r, sel


def gather_logits(local_logits: torch.Tensor, roi_index: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """All-gathers per-rank logits (variable N_r) back into the caller's ROI order.  NCCL over NVLink on GPU ranks,
    gloo in the CPU tests.  Not on the timed hot path."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    n_local = torch.tensor([local_logits.shape[0]], dtype=torch.int64, device=local_logits.device)
    sizes = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(sizes, n_local, group=group)
    sizes = [int(s.item()) for s in sizes]
    pad = max(sizes) if sizes else 0
    shape = (pad,) + tuple(local_logits.shape[1:])
    buf = torch.zeros(shape, dtype=local_logits.dtype, device=local_logits.device)
    buf[: local_logits.shape[0]] = local_logits
    idx = torch.full((pad,), -1, dtype=torch.int64, device=local_logits.device)
    idx[: roi_index.numel()] = roi_index.to(local_logits.device)
    all_buf = [torch.empty_like(buf) for _ in range(world)]
    all_idx = [torch.empty_like(idx) for _ in range(world)]
    dist.all_gather(all_buf, buf, group=group)
    dist.all_gather(all_idx, idx, group=group)
    out = torch.zeros((n_total,) + tuple(local_logits.shape[1:]), dtype=local_logits.dtype, device=local_logits.device)
    for t, i, n in zip(all_buf, all_idx, sizes):
        out[i[:n]] = t[:n]
    return out
