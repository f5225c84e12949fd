#!/usr/bin/env python
"""Benchmark of the NUNet-TLS-LSTM hot path (BASELINE.json configs[1]: offline batch = 256 synthetic 4 s clips).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch 256] [--seconds 4]

One "step" = one pass of the hot path (STFT -> network -> iSTFT/overlap-add) over one batch of `--batch` clips per
GPU.  N > 1 is launched by torchrun (one rank per GPU, NCCL); clips are independent, so ranks only share the
weight blob (one broadcast at init) and the job is weak-scaled: every rank processes its own batch.

Prints ONE JSON line (rank 0):
  value        STFT frames/s, whole job, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e          the same metric through the host entry point nunet_forward_wav_host (pinned host buffers, H2D and
               D2H inside the timed region)
  roofline     dominant kernel: algorithmic bytes / live CUDA-event duration vs the measured HBM peak;
               `path` = the whole network with SURVEY 8(d)'s B_alg = 4.256 MB/frame
  cpu_baseline the CPU oracle (restatement of the reference, torch fp32) timed on this box's host cores on a
               bounded sample of the same workload
`--impl reference` times only that CPU path (the reference's TF/TFLite runtime cannot be installed offline).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B_ALG_BYTES_PER_FRAME = 4.256e6      # SURVEY 8(d) / Appendix B, LSTM variant, offline
FLOP_PER_FRAME = 148.5e6
METRIC = "stft_frames_per_sec"
UNIT = "frames/s"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return float(j["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for r in self.rows if len(r) >= 7 and r[0].isdigit()]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(int(r[0]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": int(rows[0][1]), "samples": len(rows),
                "power_w_max": max(float(r[2]) for r in rows), "reasons": reasons}


def cpu_reference_leg(weights, n_samples: int, clips: int, steps: int, warmup: int, variant: str = "lstm"):
    """The oracle's offline forward (the reference's `model(noisy)`, test_interface.py:58) on host cores."""
    import torch
    from nunet_b200.synth import synth_clips
    from oracle.nunet_oracle import Oracle          # allowed here: the CPU baseline / reference arm
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    o = Oracle(weights, ctfa_mode="causal_avg32", variant=variant)
    wav = torch.from_numpy(synth_clips(clips, n_samples))
    T = 1 + (n_samples - 512) // 256
    with torch.no_grad():
        for _ in range(warmup):
            o.forward_wav(wav)
        t0 = time.perf_counter()
        for _ in range(steps):
            o.forward_wav(wav)
        dt = time.perf_counter() - t0
    return clips * T * steps / dt, dt / steps, cores, f"{clips} clips x {T} frames per step, offline forward, torch fp32 CPU"


B_ALG_STREAM_BYTES_PER_FRAME = 5.897e6   # 4.256 MB + 1.641 MB history read + write per stream-frame


def streaming_main(args, weights, rank, local_rank, world):
    """configs[2]: S concurrent streams, frame-basis, history carried on the device (interpreter_proposed.py:200-366
    for S streams at once).  One step = one hop of every stream."""
    import torch
    import torch.distributed as dist
    from nunet_b200.engine import NunetEngine
    from nunet_b200.synth import synth_clips
    from nunet_b200.weights import pack_blob

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    blob = pack_blob(weights)
    if world > 1:
        from nunet_b200.sharding import broadcast_blob
        dist.init_process_group("nccl", device_id=dev)
        blob = broadcast_blob(blob if rank == 0 else None, src=0, device=dev)
    S = args.streams
    eng = NunetEngine(blob, max_streams=S, device=local_rank, ctfa_mode="frame_div32", dc_mode="edge")
    nh = max(args.warmup, 4) + args.steps + 4
    pool = synth_clips(min(S, 32), 256 * nh, first_clip=1000 * rank)
    hops_h = torch.from_numpy(np.tile(pool, ((S + len(pool) - 1) // len(pool), 1))[:S]).reshape(S, nh, 256)
    hops_h = hops_h.permute(1, 0, 2).contiguous().pin_memory()          # [hop][stream][256]
    hops_d = hops_h.to(dev)
    out_d = torch.empty((S, 256), device=dev)
    out_h = torch.empty((S, 256)).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    eng.stream_reset()
    W = max(args.warmup, 4)       # 2 plain steps + one CUDA-graph capture per step parity happen before the timed region
    for i in range(W):
        eng.stream_step_wav(hops_d[i % nh], out_d)
    launches = eng.last_launch_count
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        eng.stream_step_wav(hops_d[(W + i) % nh], out_d)
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if rank == 0 else None
    value = world * S * args.steps / (ms * 1e-3)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        eng.stream_step_wav_host(hops_h[(W + i) % nh], out_h)
    torch.cuda.synchronize(dev)
    e2e = world * S * args.steps / max_over_ranks(time.perf_counter() - t0)
    if rank == 0:
        peak, peak_src = measured_peaks()
        gbs = value / world * B_ALG_STREAM_BYTES_PER_FRAME / 1e9
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": W,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"configs[2]: streaming frame-basis, {S} concurrent streams per GPU with carried conv / LSTM "
                                   "history, NUNet-TLS-LSTM, ctfa frame_div32 (one-frame graph)", "streams_per_gpu": S,
                       "l2_policy": f"per-step working set {S * B_ALG_STREAM_BYTES_PER_FRAME / 1e9:.1f} GB exceeds the 126 MB L2"},
            "rtf_per_stream": (ms / args.steps * 1e-3) / 0.016, "clocks": clocks,
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": S * 256 * 4, "d2h_bytes_per_step": S * 256 * 4},
            "gpu_launches": int(launches * args.steps),
            "roofline": {"bound": "hbm", "kernel": "whole streaming step (conv_tc3 units with the history row as a second source)", "achieved": gbs, "peak": peak,
                         "unit": "GB/s", "frac": gbs / peak, "traffic": None, "peak_source": peak_src,
                         "b_alg_bytes_per_frame": B_ALG_STREAM_BYTES_PER_FRAME},
            "cpu_baseline": None,
        }))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="clips per GPU per step")
    ap.add_argument("--seconds", type=float, default=4.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--config", default="offline", choices=["offline", "streaming"],
                    help="offline = BASELINE configs[1] (default, the headline); streaming = configs[2]: --streams "
                         "concurrent streams, one 256-sample hop per stream per step, carried conv/LSTM history")
    ap.add_argument("--streams", type=int, default=1024)
    ap.add_argument("--variant", default="lstm", choices=["lstm", "ddb"],
                    help="lstm = NUNet-TLS-LSTM (headline); ddb = configs[3], the dilated-dense baseline, offline only")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    n_samples = int(round(args.seconds * 16000))
    T = 1 + (n_samples - 512) // 256
    config = {"workload": f"configs[1]: offline batch={args.batch} x {args.seconds:g} s @16 kHz clips per GPU, "
                          f"NUNet-TLS-LSTM, T={T} frames/clip, ctfa causal_avg32",
              "clips_per_gpu": args.batch, "frames_per_clip": T, "weights": None,
              "l2_policy": "working set (inputs 65 MB + >30 GB activations per pass) far exceeds the 126 MB L2"}

    from nunet_b200.weights import (load_ddb_weights, load_default_weights, pack_blob, random_ddb_weights,
                                    random_lstm_weights)
    ddb = args.variant == "ddb"
    if ddb:
        if args.config != "offline" or args.impl != "ours":
            raise SystemExit("--variant ddb: offline config, our arm only")
        config["workload"] = (f"configs[3]: offline batch={args.batch} x {args.seconds:g} s @16 kHz clips per GPU, NUNet-TLS "
                              f"dilated-dense bottleneck variant, T={T} frames/clip, ctfa causal_avg32")
        try:
            weights = load_ddb_weights()
            config["weights"] = "shipped nutls.tflite, int8 tensors dequantised (the variant's only checkpoint)"
        except FileNotFoundError:
            weights = random_ddb_weights(0)
            config["weights"] = "RANDOM-INIT (nutls.tflite blob missing on this box)"
    else:
        try:
            weights = load_default_weights()
            config["weights"] = "trained nutls_lstm.h5 (reference checkpoint)"
        except FileNotFoundError:
            weights = random_lstm_weights(0)
            config["weights"] = "RANDOM-INIT (trained checkpoint blob missing on this box)"

    # ------------------------------------------------------------------ reference arm: CPU oracle only
    if args.impl == "reference":
        if rank != 0:
            return
        steps = max(1, min(args.steps, 3))
        val, sec, cores, sample = cpu_reference_leg(weights, n_samples, 2, steps, min(args.warmup, 1))
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": min(args.warmup, 1), "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "TensorFlow/TFLite cannot be installed offline; this is the CPU restatement of the reference path",
        }))
        return

    # ------------------------------------------------------------------ our arm
    if args.config == "streaming":
        return streaming_main(args, weights, rank, local_rank, world)
    import torch
    import torch.distributed as dist
    from nunet_b200.engine import NunetEngine
    from nunet_b200.synth import synth_clips

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    blob = pack_blob(weights, 1 if ddb else 0)
    if world > 1:
        from nunet_b200.sharding import broadcast_blob
        dist.init_process_group("nccl", device_id=dev)
        # the only collective of the path: rank 0's packed weights go to every GPU once (NVLink / NCCL)
        blob = broadcast_blob(blob if rank == 0 else None, src=0, device=dev)

    B = args.batch
    eng = NunetEngine(blob, max_frames=B * T, device=local_rank, ctfa_mode="causal_avg32", variant=1 if ddb else 0)
    b_alg = 4.358e6 if ddb else B_ALG_BYTES_PER_FRAME
    flop = 149.0e6 if ddb else FLOP_PER_FRAME
    # synthetic clips: a pool of 32 distinct clips tiled to the batch (generation is host-side numpy)
    pool = synth_clips(min(B, 32), n_samples, first_clip=1000 * rank)
    wav_h = torch.from_numpy(np.tile(pool, ((B + len(pool) - 1) // len(pool), 1))[:B]).contiguous().pin_memory()
    wav_d = wav_h.to(dev)
    out_d = torch.empty((B, (T - 1) * 256 + 512), device=dev, dtype=torch.float32)
    out_h = torch.empty(out_d.shape, dtype=torch.float32).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput
    for _ in range(max(args.warmup, 3)):
        eng.forward_wav_into(wav_d, out_d)
    launches_per_step = eng.last_launch_count
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        eng.forward_wav_into(wav_d, out_d)
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if rank == 0 else None
    frames_total = world * B * T * args.steps
    value = frames_total / (ms * 1e-3)

    # ---- end to end through the host entry point (H2D + kernels + D2H per step)
    for _ in range(2):
        eng.forward_wav_host(wav_h, out_h)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        eng.forward_wav_host(wav_h, out_h)
    torch.cuda.synchronize(dev)
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = frames_total / e2e_s

    # ---- per-kernel profile (one extra step, events after every launch) -> dominant kernel roofline
    roofline = None
    if rank == 0:
        peak, peak_src = measured_peaks()
        eng.profile(True)
        eng.forward_wav_into(wav_d, out_d)
        torch.cuda.synchronize(dev)
        ent = eng.profile_entries()
        eng.profile(False)
        total_ms = sum(e[1] for e in ent)
        by_kernel = {}
        for name, kms, nbytes in ent:
            k = by_kernel.setdefault(name, [0.0, 0.0])
            k[0] += kms
            k[1] += nbytes
        top_name, (top_ms, top_bytes) = max(by_kernel.items(), key=lambda kv: kv[1][0])
        achieved = top_bytes / (top_ms * 1e-3) / 1e9
        # DRAM bytes of that launch from the committed ncu capture of the same command line (profiles/), if the
        # capture was taken at this batch size
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "r1_tc3_dram_traffic.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            u = tj["units"].get(top_name)
            if u is not None and tj["frames"] == B * T:
                traffic = u["dram_read_bytes"] + u["dram_write_bytes"]
        path_gbs = value / world * b_alg / 1e9
        roofline = {"bound": "hbm", "kernel": top_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "algorithmic_bytes": top_bytes,
                    "peak_source": peak_src,
                    "kernel_ms": top_ms, "kernel_share_of_step": top_ms / total_ms,
                    "path": {"achieved": path_gbs, "frac": path_gbs / peak,
                             "b_alg_bytes_per_frame": b_alg,
                             "fp32_tflops": value / world * flop / 1e12},
                    "top5": sorted(((n, round(v[0], 3)) for n, v in by_kernel.items()), key=lambda x: -x[1])[:5]}
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "bench_kernel_profile.json"), "w") as f:
            json.dump({"total_ms": total_ms, "entries": ent}, f)

    cpu_baseline = None
    if rank == 0 and not args.no_cpu_baseline:
        val, sec, cores, sample = cpu_reference_leg(weights, n_samples, 2, 2, 1, variant=args.variant)
        cpu_baseline = {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}

    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "rtf_per_stream_equivalent": (1.0 / (value / world)) / 0.016,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(wav_h.numel() * 4),
                    "d2h_bytes_per_step": int(out_h.numel() * 4)},
            "gpu_launches": int(launches_per_step * args.steps),
            "roofline": roofline, "cpu_baseline": cpu_baseline,
        }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
