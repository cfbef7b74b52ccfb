#!/usr/bin/env python
"""bench.py -- Diff-UNet DDIM-10 sliding-window inference throughput (BASELINE.json metric: 96^3 patches/s).

    python bench.py --gpus N --steps K --warmup W            # B200 path (this repo)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port) on the host cores

One "step" = one pass of the hot path over one synthetic 512x512x160 CT volume (AMOS, C=16, roi 96^3, overlap 0.25,
98 windows, DDIM-10).  For N > 1 the 98 windows are sharded contiguously over the ranks (no data-path collective until
the single NCCL reduce of the stitched logits to rank 0); per-GPU work shrinks as N grows -> "strong" scaling.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

VOLUME = (512, 512, 160)
ROI = (96, 96, 96)
CLASSES = 16
FEATURES = (64, 64, 128, 256, 512, 64)
OVERLAP = 0.25
STEPS_DDIM = 10
METRIC = "96^3 patches/s (DDIM-10)"
WORKLOAD = "AMOS 16-class sliding-window DDIM-10 inference, synthetic 512x512x160 CT volume, roi 96^3, overlap 0.25 (98 windows)"

# algorithmic FLOPs of the path: encoder 279.8 GFLOP/window + 10 x 1056.4 GFLOP
GFLOP_PER_WINDOW = 279.8 + 10 * 1056.4


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi SM clocks / throttle reasons sampled every 200 ms during the timed region."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port of the reference's CPU path on the host cores
# ----------------------------------------------------------------------------------------------------------------
def cpu_window_sample(threads: int):
    """Bounded sample of one 96^3 window of the reference path: the encoder once + ONE of the ten denoiser/DDIM steps,
    fp32, torch CPU.  Returns (seconds_encoder, seconds_step)."""
    import torch

    from oracle import oracle_ddim, oracle_model

    torch.set_num_threads(threads)
    sd = oracle_model.init_state_dict(1, CLASSES, FEATURES, seed=0)
    torch.manual_seed(1)
    image = torch.rand(1, 1, *ROI)
    torch.manual_seed(2)
    x = torch.randn(1, CLASSES, *ROI)
    sched = oracle_ddim.SpacedSchedule(STEPS_DDIM)
    with torch.no_grad():
        t0 = time.perf_counter()
        emb = oracle_model.encoder_forward(sd, image)
        t1 = time.perf_counter()
        out = oracle_model.denoiser_forward(sd, x, torch.tensor([sched.timestep_map[-1]]), image, emb)
        oracle_ddim.ddim_step(sched, sched.num_timesteps - 1, x, out)
        t2 = time.perf_counter()
    return t1 - t0, t2 - t1


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    cpu_window_sample(threads) if args.warmup > 0 else None  # one warm-up sample (allocator, oneDNN primitives)
    t_all, vals = time.perf_counter(), []
    for _ in range(max(args.steps, 1)):
        te, ts = cpu_window_sample(threads)
        vals.append(1.0 / (te + STEPS_DDIM * ts))
    elapsed = time.perf_counter() - t_all
    v = sum(vals) / len(vals)
    sample = "per step: encoder + 1 of 10 DDIM steps of one 96^3 window (C=16, fp32 torch CPU); patches/s = 1/(t_enc + 10*t_step)"
    print_json({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "patches/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * elapsed / max(args.steps, 1), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "reference CPU path (oracle port of the reference's PyTorch code; the reference is pure Python and cannot travel to the GPU box)"},
        "cpu_baseline": {"value": v, "unit": "patches/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "patches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


# ----------------------------------------------------------------------------------------------------------------
# B200 arm
# ----------------------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist

    import diff_unet_amos_b200 as pkg
    from diff_unet_amos_b200 import _lib
    from diff_unet_amos_b200.inference import StitchBuffers, crop_windows

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = pkg.load_library()

    torch.manual_seed(0)
    model = pkg.DiffUNetB200(in_channels=1, out_channels=CLASSES, image_size=ROI[1], spatial_size=ROI[0], features=FEATURES,
                             batch_max=args.sw_batch + 1, precision=args.precision,
                             dual_stream=bool(args.dual_stream)).to(dev).eval()
    torch.manual_seed(1)
    host_vol = torch.rand(1, 1, *VOLUME).pin_memory()
    dev_vol = host_vol.to(dev)
    starts = pkg.window_starts(VOLUME, ROI, args.overlap)
    n_win = len(starts)
    lo, hi = pkg.shard_range(n_win, rank, world)
    gen = torch.Generator(device=dev)
    gen.manual_seed(2 + rank)
    host_labels = torch.empty((CLASSES,) + VOLUME, dtype=torch.uint8).pin_memory() if rank == 0 else None
    # second pinned result buffer for the overlapped D2H of the e2e leg -- allocated here: cudaHostAlloc of 671 MB takes
    # ~0.5 s and must not sit between the barrier and the timed region of rank 0 (the other ranks would wait for it)
    host_labels2 = [host_labels, torch.empty((CLASSES,) + VOLUME, dtype=torch.uint8).pin_memory()] if rank == 0 else None

    # window batches of this rank: sw_batch windows each; a single left-over window joins the last batch (13 = 4 + 4 + 5)
    # instead of running alone at batch-1 efficiency
    bounds = list(range(lo, hi, args.sw_batch)) + [hi]
    if len(bounds) > 2 and bounds[-1] - bounds[-2] == 1:
        del bounds[-2]
    scatter = world > 1 and CLASSES % world == 0

    def one_volume(volume_dev):
        """the hot path for this rank's shard of windows, then the NCCL exchange (reduce-scatter of the stitched logits by
        channel, local divide + binarise, gather of the uint8 labels on rank 0)"""
        buf = StitchBuffers(CLASSES, VOLUME, ROI, args.overlap, dev)
        for g0, g1 in zip(bounds[:-1], bounds[1:]):
            grp = starts[g0:g1]
            batch = crop_windows(volume_dev[0], grp, ROI)
            noise = torch.randn((len(grp), CLASSES) + ROI, device=dev, generator=gen)  # gaussian_diffusion.py:693
            pred = model(image=batch, pred_type="ddim_sample", noise=noise)
            for j, s in enumerate(grp):
                buf.add(pred[j], s)
        if scatter:
            mine = StitchBuffers.__new__(StitchBuffers)
            mine.vol, mine.roi, mine.mode, mine.counts, mine.channels = buf.vol, buf.roi, buf.mode, buf.counts, CLASSES // world
            mine.out = pkg.reduce_scatter_channels(buf.out)      # sum of the partial volumes, my channels only
            return pkg.gather_channel_chunks(mine.finalize(binary=True)[1], dst=0)
        if world > 1:
            dist.reduce(buf.out, dst=0)
        if rank == 0:
            return buf.finalize(binary=True)[1]
        return None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        for _ in range(args.warmup):
            one_volume(dev_vol)
        # ---------------- device-resident leg: `value` ----------------
        _lib.check(lib.dunet_profile_enable(1))
        barrier()
        clocks = ClockSampler(local) if rank == 0 else None
        l0 = lib.dunet_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            one_volume(dev_vol)
        e1.record()
        barrier()
        launches = lib.dunet_launch_count() - l0
        clk = clocks.stop() if clocks else None
        ms = e0.elapsed_time(e1)
        conv_ms, conv_n, conv_fl = ctypes.c_double(), ctypes.c_uint64(), ctypes.c_double()
        _lib.check(lib.dunet_profile_read(ctypes.byref(conv_ms), ctypes.byref(conv_n), ctypes.byref(conv_fl)))
        fam_ms, fam_n, fam_b = (ctypes.c_double * 8)(), (ctypes.c_uint64 * 8)(), (ctypes.c_double * 8)()
        _lib.check(lib.dunet_profile_read_all(fam_ms, fam_n, fam_b))
        _lib.check(lib.dunet_profile_enable(0))
        # ---------------- end-to-end leg: host volume in, host labels out, every step ----------------
        # Both copies of every step are inside the timed region.  They run on a side stream so that the D2H of volume i's
        # labels overlaps the windows of volume i + 1 (two pinned label buffers); the last D2H is waited for before the
        # closing event.
        barrier()
        copy_stream, d2h_stream = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        main = torch.cuda.current_stream()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for it in range(args.steps):
            with torch.cuda.stream(copy_stream):
                v = host_vol.to(dev, non_blocking=True)      # H2D of the step's input from pinned memory
            main.wait_stream(copy_stream)
            v.record_stream(main)
            lab = one_volume(v)
            if rank == 0:
                d2h_stream.wait_stream(main)
                with torch.cuda.stream(d2h_stream):
                    host_labels2[it % 2].copy_(lab, non_blocking=True)  # D2H of the step's result (binary label volume)
                lab.record_stream(d2h_stream)
        main.wait_stream(d2h_stream)
        f1.record()
        barrier()
        ms_e2e = f0.elapsed_time(f1)
    if world > 1:
        t = torch.tensor([ms, ms_e2e, float(launches)], device=dev, dtype=torch.float64)
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        ms, ms_e2e, launches = float(tmax[0]), float(tmax[1]), int(t[2])
    if rank == 0:
        peaks = measured_peaks()
        value = n_win * args.steps / (ms / 1e3)
        e2e = n_win * args.steps / (ms_e2e / 1e3)
        conv_tflops = conv_fl.value / (conv_ms.value / 1e3) / 1e12 if conv_ms.value > 0 else 0.0
        peak = peaks["bf16_tflops_sustained"]
        out = {
            "metric": METRIC, "value": value, "unit": "patches/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "bf16x3 (fp32-class split operands)", "data": "synthetic",
            "config": {"workload": WORKLOAD if args.overlap == OVERLAP else WORKLOAD.replace("overlap 0.25 (98 windows)", f"overlap {args.overlap} ({n_win} windows)"), "features": list(FEATURES), "classes": CLASSES, "sw_batch": args.sw_batch,
                       "windows_per_step": n_win, "windows_this_rank": hi - lo, "volumes_per_s": value / n_win,
                       "dual_stream": "e2e leg only (half batches on two internal streams; off while per-kernel profiling is on)" if args.dual_stream else "off",
                       "l2": "inputs larger than L2 (each window streams > 1 GB of activations; no flush needed)",
                       "algorithmic_tflop_per_window": GFLOP_PER_WINDOW / 1e3,
                       "whole_path_tflops": value * GFLOP_PER_WINDOW / 1e3,
                       "whole_path_frac_of_bf16_peak": value * GFLOP_PER_WINDOW / 1e3 / world / peak},
            "clocks": clk,
            "e2e": {"value": e2e, "unit": "patches/s", "h2d_bytes_per_step": host_vol.numel() * 4,
                    "d2h_bytes_per_step": host_labels.numel(), "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "kernel": "conv3d_tc64_kernel + conv3d_tc_kernel: every 3x3x3 conv launch of rank 0 in the timed region (CUDA events around each launch)",
                         "achieved": conv_tflops, "peak": peak, "unit": "TFLOP/s", "frac": conv_tflops / peak,
                         "peak_source": peaks["source"] + ", sustained cuBLAS bf16",
                         "launches": int(conv_n.value), "avg_launch_ms": conv_ms.value / max(conv_n.value, 1),
                         "conv_share_of_step": conv_ms.value / ms,
                         "frac_of_burst_peak": conv_tflops / peaks["bf16_tflops"],
                         "traffic": 864.9e6, "traffic_note": "ncu --set full, conv3d_tc64_kernel<32,4,0> 64->64 @96^3, 4 windows per launch: dram read 453.4 MB + write 411.5 MB (algorithmic 4 x 113.2 MB each way; part of the output is still in L2 at kernel end), tensor pipe active 79 % of elapsed (profiles/r1_ncu_full_conv3d_tc64_batch4_v2.txt)"},
            "roofline_hbm": {name: {"bound": "hbm", "achieved": fam_b[i] / fam_ms[i] / 1e6, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                    "frac": fam_b[i] / fam_ms[i] / 1e6 / peaks["hbm_gbs"], "launches": int(fam_n[i]),
                                    "share_of_step": fam_ms[i] / ms}
                             for i, name in ((1, "normalise (IN+LeakyReLU+bias+residual+pool), launches >= 64 MB"), (2, "final 1x1 conv + DDIM update + accumulate"),
                                             (3, "transposed conv k2s2"), (6, "normalise, launches < 64 MB (launch-latency bound)")) if fam_ms[i] > 0},
            "kernel_time_share": {n: fam_ms[i] / ms for i, n in enumerate(["conv3x3x3", "normalise_large", "final_ddim", "deconv", "splitk_reduce", "affine_map", "normalise_small"])},
        }
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            te, ts = cpu_window_sample(threads)
            out["cpu_baseline"] = {"value": 1.0 / (te + STEPS_DDIM * ts), "unit": "patches/s", "cores": threads, "kind": "port",
                                   "sample": "encoder + 1 of 10 DDIM steps of one 96^3 window, fp32 torch CPU oracle port; patches/s = 1/(t_enc + 10*t_step)"}
        print_json(out)
    if world > 1:
        dist.destroy_process_group()


def _claim_stdout():
    """Libraries (NCCL prints its version banner) write to fd 1; the contract is ONE JSON line on stdout.  Route fd 1 to
    stderr for the duration of the run and return a writer bound to the real stdout."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(real, "w")


def main():
    out_stream = _claim_stdout()
    global print_json
    def print_json(obj):
        out_stream.write(json.dumps(obj) + "\n")
        out_stream.flush()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--sw-batch", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--overlap", type=float, default=OVERLAP, help="0.25 = test.py:30 default (98 windows); 0.8 = cfg/btcv, cfg/msd (2645 windows)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32x3"])
    ap.add_argument("--dual-stream", type=int, default=1, help="DUNET_FLAG_DUAL_STREAM: two half batches on two internal streams (the product default; "
                    "inactive in the profiled device-resident leg)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
