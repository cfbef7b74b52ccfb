"""world_size-2 gloo test (CPU) of the multi-GPU host logic: contiguous window shards + one sum-reduce of the partial
stitched volumes reproduces the single-process MONAI-order result."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle_sliding


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_pred(image, start, roi):
    """integer-valued 'prediction' of a window so that fp32 sums are exact in any order"""
    z, y, x = start
    crop = image[0, 0, z:z + roi[0], y:y + roi[1], x:x + roi[2]]
    return torch.stack([torch.round(crop * 8), torch.full_like(crop, float(z + 2 * y + 3 * x))])


def _worker(rank, world, port, vol, roi, overlap, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from diff_unet_amos_b200 import (gather_channel_chunks, my_window_range, reduce_partial_volume, reduce_scatter_channels,
                                     window_starts)

    torch.manual_seed(1)
    image = torch.rand(1, 1, *vol)
    starts = window_starts(vol, roi, overlap)
    lo, hi = my_window_range(len(starts))
    part = torch.zeros((2,) + tuple(vol))
    for s in starts[lo:hi]:
        part[:, s[0]:s[0] + roi[0], s[1]:s[1] + roi[1], s[2]:s[2] + roi[2]] += _fake_pred(image, s, roi)
    # path 2: reduce-scatter by channel + gather of the chunks (what the NCCL path does, so that every rank finalizes
    # its own channels) must give the same sum
    chunk = reduce_scatter_channels(part.clone())
    assert chunk.shape[0] == part.shape[0] // world
    full = gather_channel_chunks(chunk, dst=0)
    reduce_partial_volume(part, dst=0)
    if rank == 0:
        ret["sum"] = part.numpy().copy()
        ret["sum_scatter"] = full.numpy().copy()
        ret["ranges"] = [my_window_range(len(starts), r, world) for r in range(world)]
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_window_sharding_matches_single_process():
    vol, roi, overlap = (48, 40, 36), (32, 32, 32), 0.25
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), vol, roi, overlap, ret), nprocs=2, join=True)
    torch.manual_seed(1)
    image = torch.rand(1, 1, *vol)
    grid = oracle_sliding.window_grid(vol, roi, overlap)
    ref = torch.zeros((2,) + vol)
    for s in grid:
        ref[:, s[0]:s[0] + roi[0], s[1]:s[1] + roi[1], s[2]:s[2] + roi[2]] += _fake_pred(image, s, roi)
    assert np.array_equal(ret["sum"], ref.numpy())
    assert np.array_equal(ret["sum_scatter"], ref.numpy())
    (lo0, hi0), (lo1, hi1) = ret["ranges"]
    assert lo0 == 0 and hi0 == lo1 and hi1 == len(grid) and abs((hi0 - lo0) - (hi1 - lo1)) <= 1
