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
    outs = []
    for b in bufs:
        reduce_partial_volume(b.out, dst)
        if not dist.is_initialized() or dist.get_rank() == dst:
            outs.append(b.finalize(binary=True))
    if dist.is_initialized() and dist.get_rank() != dst:
        return None, None
    blended = torch.stack([o[0] for o in outs])
    labels = torch.stack([o[1] for o in outs]).float()
    return blended, labels
