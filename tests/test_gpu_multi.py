"""2-GPU NCCL test of infer_volume_distributed / infer_volumes_distributed (skipped with fewer than 2 GPUs): window shards +
reduce-scatter by channel + local finalize + gather == the single-GPU result, incl. a volume smaller than the roi (padding must
be cropped back) and the throughput-mode window queues."""
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

    cout, S = 4, 32
    torch.manual_seed(0)
    m = pkg.DiffUNetB200(in_channels=1, out_channels=cout, image_size=S, spatial_size=S, features=SMALL).cuda().eval()
    # (48, 56, 40): plain case.  (48, 56, 24): one axis smaller than the roi -> the driver pads and must crop back
    for tag, vol in (("", (48, 56, 40)), ("_padded", (48, 56, 24))):
        image = seeded_image((1, 1) + vol).cuda()
        padded = tuple(max(v, S) for v in vol)
        n_win = len(pkg.window_starts(padded, (S, S, S), 0.25))
        noise = seeded_noise((n_win, cout, S, S, S)).cuda()
        nf = lambda w, b: noise[w:w + b]
        blended, labels = pkg.infer_volume_distributed(m, image, sw_batch_size=2, overlap=0.25, noise_fn=nf)
        if rank == 0:
            ref_b, ref_l = pkg.infer_volume(m, image, sw_batch_size=2, overlap=0.25, noise_fn=nf)
            ret["rel" + tag] = float((blended - ref_b).norm() / ref_b.norm())
            ret["agree" + tag] = float((labels == ref_l).float().mean())
            ret["shape_ok" + tag] = (tuple(blended.shape) == tuple(ref_b.shape) == (1, cout) + vol
                                      and tuple(labels.shape) == tuple(ref_l.shape))
        else:
            assert blended is None and labels is None
    # throughput mode: three volumes as window queues; library noise keyed by (seed, global window index) -> the same
    # volumes one by one on a single GPU give the same labels
    vol = (48, 56, 40)
    images = [seeded_image((1, 1) + vol, seed=20 + i).cuda() for i in range(3)]
    n_win = len(pkg.window_starts(vol, (S, S, S), 0.25))
    outs = pkg.infer_volumes_distributed(m, images, sw_batch_size=2, overlap=0.25, seed=11)
    if rank == 0:
        agree = []
        for i, img in enumerate(images):
            bufs = pkg.sliding_window_inference(img, (S, S, S), 2, m, 0.25, finalize=False, pred_type="ddim_sample",
                                                noise_fn=None, seed=11)
            # single-GPU numbering of volume i's windows in the queue: i * n_win + w
            b = pkg.StitchBuffers(cout, vol, (S, S, S), 0.25, "cuda")
            st = pkg.window_starts(vol, (S, S, S), 0.25)
            for g in range(0, n_win, 2):
                b.add_windows(m, img[0, 0], st[g:g + 2], seed=11, noise_ids=range(i * n_win + g, i * n_win + min(g + 2, n_win)))
            ref = b.finalize(binary=True)[1]
            agree.append(float((outs[i] == ref).float().mean()))
        ret["queue_agree"] = min(agree)
        ret["queue_n"] = len(outs)
    else:
        assert outs is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_volume_equals_single_gpu():
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    for tag in ("", "_padded"):
        assert ret["shape_ok" + tag], tag
        assert ret["rel" + tag] < 1e-6 and ret["agree" + tag] > 0.9999, tag  # fp32 sums in a different order only
    assert ret["queue_n"] == 3 and ret["queue_agree"] > 0.9999
    print("2-GPU check:", dict(ret))
