"""2-GPU NCCL test of infer_volume_distributed (skipped with fewer than 2 GPUs): window shards + reduce-scatter by channel
+ local finalize + gather == the single-GPU result."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.util import SMALL, seeded_image, seeded_noise

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import diff_unet_amos_b200 as pkg

    cout, S, vol = 4, 32, (48, 56, 40)
    torch.manual_seed(0)
    m = pkg.DiffUNetB200(in_channels=1, out_channels=cout, image_size=S, spatial_size=S, features=SMALL).cuda().eval()
    image = seeded_image((1, 1) + vol).cuda()
    n_win = len(pkg.window_starts(vol, (S, S, S), 0.25))
    noise = seeded_noise((n_win, cout, S, S, S)).cuda()
    nf = lambda w, b: noise[w:w + b]
    blended, labels = pkg.infer_volume_distributed(m, image, sw_batch_size=2, overlap=0.25, noise_fn=nf)
    if rank == 0:
        ref_b, ref_l = pkg.infer_volume(m, image, sw_batch_size=2, overlap=0.25, noise_fn=nf)
        ret["rel"] = float((blended - ref_b).norm() / ref_b.norm())
        ret["agree"] = float((labels == ref_l).float().mean())
        ret["shape_ok"] = tuple(blended.shape) == tuple(ref_b.shape) and tuple(labels.shape) == tuple(ref_l.shape)
    else:
        assert blended is None and labels is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_volume_equals_single_gpu():
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    assert ret["shape_ok"]
    assert ret["rel"] < 1e-6 and ret["agree"] > 0.9999  # fp32 sums in a different order only
