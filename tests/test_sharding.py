"""CPU: the N>1 host logic (image-range partitioning, ROI rebasing, result gather) with world_size-2 gloo."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from human_instance_segmentation_b200 import sharding
from tests import common


def test_partition_is_contiguous_balanced_and_complete():
    rois = common.synth_rois(3, 16, 1)
    extra = common.synth_rois(4, 16, 9)[torch.arange(0, 144, 3)]          # ragged: some images get many ROIs
    rois = torch.cat([rois, extra], 0)
    for world in (1, 2, 4, 8):
        bounds = sharding.partition_images(rois[:, 0], 16, world)
        assert bounds[0][0] == 0 and bounds[-1][1] == 16
        assert all(bounds[i][1] == bounds[i + 1][0] for i in range(world - 1))
        assert all(hi > lo for lo, hi in bounds)
        seen = []
        for r in range(world):
            im, rr, idx = sharding.shard_batch(torch.zeros(16, 3, 4, 4), rois, world, r, bounds)
            assert im.shape[0] == bounds[r][1] - bounds[r][0]
            assert rr.shape[0] == 0 or (0 <= rr[:, 0].min() and rr[:, 0].max() < im.shape[0])
            assert torch.equal(rr[:, 1:], rois[idx][:, 1:])
            seen.append(idx)
        assert torch.equal(torch.cat(seen).sort()[0], torch.arange(rois.shape[0]))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rois = torch.cat([common.synth_rois(8, 6, 2), common.synth_rois(9, 6, 1)], 0)
    images = torch.arange(6, dtype=torch.float32).view(6, 1, 1, 1).expand(6, 3, 2, 2).contiguous()
    im, rr, idx = sharding.shard_batch(images, rois, world, rank)
    # stand-in for the per-rank forward: "logits" that encode (original image id, x1) of every ROI
    lo = int(im[0, 0, 0, 0].item())
    local = torch.stack([rr[:, 0] + lo, rr[:, 1]], 1).view(-1, 2, 1, 1)
    full = sharding.gather_logits(local, idx, rois.shape[0])
    ok = torch.equal(full[:, 0, 0, 0], rois[:, 0]) and torch.equal(full[:, 1, 0, 0], rois[:, 1])
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_shard_and_gather_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
    assert sorted(res) == [(0, True), (1, True)]
