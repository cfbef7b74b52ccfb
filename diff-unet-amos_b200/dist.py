"""Multi-GPU sliding-window inference: one process per GPU, windows sharded contiguously, ONE collective.

The reference runs inference on a single GPU (engine.py:172-177); its only distributed idiom on the evaluation side is a
contiguous per-rank shard followed by a gather (light_training/sampler.py:5-48).  Windows are independent units, so the
same idea applies: rank r processes windows [lo, hi) of MONAI's window order into its own partial fp32 sum volume, the
partial volumes are summed onto rank 0 (NCCL reduce over NVLink: the "gather of the stitched logits"), and rank 0
divides by the analytically known coverage counts.  No collective touches the per-window data path.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist

from .windows import shard_range, window_starts


def my_window_range(n_windows: int, rank: Optional[int] = None, world: Optional[int] = None) -> Tuple[int, int]:
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    return shard_range(n_windows, rank, world)


def reduce_partial_volume(partial: torch.Tensor, dst: int = 0) -> torch.Tensor:
    """Sum the per-rank partial stitched volumes onto ``dst`` (in place on ``dst``).  Works with NCCL (CUDA tensors) and
    gloo (CPU tensors, used by the CPU tests)."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(partial, dst=dst, op=dist.ReduceOp.SUM)
    return partial


def reduce_scatter_channels(partial: torch.Tensor) -> torch.Tensor:
    """Sum the per-rank partial stitched volumes [C, D, H, W] and leave rank r with channels [r*C/N, (r+1)*C/N) of the
    sum (C must be divisible by the world size).  NCCL: one reduce-scatter over NVLink -- every rank then divides /
    binarises its own channels, so the fp32 volume is never gathered.  gloo (CPU tests) has no reduce-scatter: all-reduce
    then slice."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return partial
    C = partial.shape[0]
    if C % world:
        raise ValueError(f"{C} channels cannot be split over {world} ranks")
    per = C // world
    rank = dist.get_rank()
    if partial.is_cuda:
        out = torch.empty((per,) + tuple(partial.shape[1:]), dtype=partial.dtype, device=partial.device)
        dist.reduce_scatter_tensor(out, partial.contiguous(), op=dist.ReduceOp.SUM)
        return out
    dist.all_reduce(partial, op=dist.ReduceOp.SUM)
    return partial[rank * per:(rank + 1) * per].clone()


def gather_channel_chunks(chunk: torch.Tensor, dst: int = 0):
    """Concatenate every rank's channel chunk (in rank order) on ``dst``; returns None elsewhere."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return chunk
    chunk = chunk.contiguous()
    if dist.get_rank() == dst:
        parts = [torch.empty_like(chunk) for _ in range(world)]
        dist.gather(chunk, gather_list=parts, dst=dst)
        return torch.cat(parts, dim=0)
    dist.gather(chunk, gather_list=None, dst=dst)
    return None


@torch.no_grad()
def infer_volume_distributed(model, image: torch.Tensor, sw_batch_size: int = 4, overlap: float = 0.25,
                             noise_fn: Optional[Callable[[int, int], torch.Tensor]] = None, dst: int = 0):
    """Engine.infer (engine.py:167-182) with the window list sharded over the process group.

    Every rank holds the full input volume and the same weights.  Returns (blended volume, binary labels) on ``dst``
    and (None, None) elsewhere.  ``noise_fn(first_window_index, count)`` supplies explicit noise (parity runs)."""
    from .inference import sliding_window_inference

    roi = model.patch
    vol = tuple(max(int(i), int(r)) for i, r in zip(image.shape[2:], roi))
    n_win = len(window_starts(vol, roi, overlap))
    lo, hi = my_window_range(n_win)
    cursor = {"w": lo}

    def predictor(batch, pred_type=None):
        nz = noise_fn(cursor["w"], batch.shape[0]) if noise_fn is not None else None
        cursor["w"] += batch.shape[0]
        return model(image=batch, pred_type=pred_type, noise=nz)

    bufs = sliding_window_inference(image, roi, sw_batch_size, predictor, overlap, window_range=(lo, hi), finalize=False,
                                    out_channels=model.num_classes, pred_type="ddim_sample")
    world = dist.get_world_size() if dist.is_initialized() else 1
    outs = []
    for b in bufs:
        if world > 1 and b.channels % world == 0 and b.mode == "constant":
            # reduce-scatter by channel, finalize locally (the count map does not depend on the channel), gather results
            from .inference import StitchBuffers

            mine = StitchBuffers.__new__(StitchBuffers)
            mine.vol, mine.roi, mine.mode, mine.counts = b.vol, b.roi, b.mode, b.counts
            mine.channels = b.channels // world
            mine.out = reduce_scatter_channels(b.out)
            blended_c, binary_c, _ = mine.finalize(binary=True)
            blended, binary = gather_channel_chunks(blended_c, dst), gather_channel_chunks(binary_c, dst)
            if dist.get_rank() == dst:
                outs.append((blended, binary))
        else:
            reduce_partial_volume(b.out, dst)
            if not dist.is_initialized() or dist.get_rank() == dst:
                outs.append(b.finalize(binary=True)[:2])
    if dist.is_initialized() and dist.get_rank() != dst:
        return None, None
    blended = torch.stack([o[0] for o in outs])
    labels = torch.stack([o[1] for o in outs]).float()
    return blended, labels
