#!/usr/bin/env python
"""Benchmark of the NUNet-TLS hot path (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--config all|offline|streaming] [--variant lstm|ddb] [--batch 256] [--seconds 4] [--streams 1024]

Headline (the one JSON line rank 0 prints) = BASELINE configs[1]: offline batch of `--batch` synthetic 4 s clips per
GPU through NUNet-TLS-LSTM.  One "step" = one pass of the hot path (STFT -> network -> iSTFT / overlap-add) over one
batch.  N > 1 is launched by torchrun (one rank per GPU, NCCL); clips are independent, so ranks only share the weight
blob (one broadcast at init) and the job is weak-scaled: every rank processes its own batch.

With `--config all` (the default) and one GPU the same line carries, under "extra", the records of
  configs[2]  streaming, `--streams` concurrent streams, one 256-sample hop of every stream per step
  configs[3]  the dilated-dense (DDB) variant, offline, same batch
each with its own value / e2e / roofline / launch count, measured in the same process right after the headline.

Keys of a record:
  value        STFT frames/s, whole job, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e          the same metric through the host entry point (pinned host buffers, H2D and D2H inside the timed region)
  roofline     the PATH against the measured HBM peak: achieved = frames/s x B_alg (algorithmic bytes per
               frame), `traffic` = DRAM bytes per step when an ncu capture of this workload is committed under
               profiles/ (named in `traffic_source`), else null; `dominant_kernel` = the same figure for the single
               kernel (one launch) with the largest share of the step, timed live with CUDA events after every launch
  cpu_baseline the CPU oracle (restatement of the reference, torch fp32) timed on this box's host cores: offline form on
               a bounded sample, and the reference's own form -- frame by frame, batch 1 (interpreter_proposed.py:
               200-366) -- at 1 thread and at all cores, with ms/frame and RTF (:411-412)
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

# SURVEY 8(d) / Appendix B
B_ALG = {"lstm": 4.256e6, "ddb": 4.358e6}                     # offline algorithmic bytes per frame
B_ALG_STREAM = {"lstm": 4.256e6 + 1.641e6, "ddb": 4.358e6 + 3.295e6}   # + history read + write per stream-frame
FLOP_PER_FRAME = {"lstm": 148.5e6, "ddb": 149.0e6}
METRIC = "stft_frames_per_sec"
UNIT = "frames/s"
DTYPE = "f32"            # fp32 in / out and fp32 accumulation; the conv contractions run as 3 fp16 hi/lo products on tcgen05
DTYPE_NOTE = "f32 results: conv contractions as split-half fp16 hi/lo x3 on tcgen05 with fp32 accumulation (parity-equivalent to fp32)"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return float(j["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_tensor_peak():
    """sustained dense bf16 / fp16 TFLOP/s (the step is long: the sustained figure applies)"""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return float(j.get("bf16_tflops_sustained", j.get("bf16_tflops", 1410.0))), "measured (MEASURED_PEAKS.json, sustained)"
    return 1410.0, "fallback (B200_PROFILING.md)"


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


# ------------------------------------------------------------------------------------------------ CPU legs (oracle)
def cpu_offline_leg(weights, n_samples: int, clips: int, steps: int, warmup: int, variant: str = "lstm"):
    """The oracle's offline forward (the reference's `model(noisy)`, test_interface.py:58) on all host cores."""
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
    return clips * T * steps / dt, dt / steps, cores, (f"{clips} clips x {T} frames per step x {steps} steps, offline forward "
                                                      f"(test_interface.py:58 form), torch fp32 CPU, {cores} threads")


def cpu_streaming_leg(weights, frames: int, threads: int, variant: str = "lstm"):
    """The reference's OWN form of the path: `real_time_speech_enhancer` -- batch 1, one signature call per 256-sample hop
    with the whole history dict round-tripped, numpy framing around it (interpreter_proposed.py:200-366) -- timed like the
    reference does (`time_array` per frame :201,366; RTF = mean / 16 ms :411-412)."""
    import torch
    from nunet_b200.synth import synth_clips
    from oracle.nunet_oracle import Oracle
    torch.set_num_threads(threads)
    o = Oracle(weights, ctfa_mode="frame_div32", variant=variant)
    wav = synth_clips(1, 256 * (frames + 4))[0]
    _, times = o.real_time_speech_enhancer(wav, dc_pad="edge")
    times = np.asarray(times[3:])                      # first calls pay lazy initialisation
    ms = float(times.mean() * 1e3)
    return {"threads": threads, "frames": int(len(times)), "ms_per_frame": ms, "frames_per_s": 1e3 / ms, "rtf": ms / 16.0}


def cpu_baseline_record(weights, n_samples, variant="lstm", offline_steps=2, stream_frames=60):
    val, sec, cores, sample = cpu_offline_leg(weights, n_samples, 2, offline_steps, 1, variant=variant)
    rec = {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
           "label": "restated CPU baseline (TF/TFLite unavailable offline)"}
    try:
        rec["streaming_form"] = {
            "what": "oracle frame loop, batch 1, one frame per call (interpreter_proposed.py:200-366), one synthetic clip",
            "k1": cpu_streaming_leg(weights, stream_frames, 1, variant),
            "kall": cpu_streaming_leg(weights, stream_frames, cores, variant)}
    except Exception as e:                              # the offline figure stands on its own
        rec["streaming_form"] = {"error": repr(e)}
    return rec


# ------------------------------------------------------------------------------------------------ GPU legs
class Ctx:
    def __init__(self, rank, local_rank, world):
        import torch
        self.rank, self.local_rank, self.world = rank, local_rank, world
        self.dev = torch.device("cuda", local_rank)

    def barrier(self):
        import torch
        import torch.distributed as dist
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, x: float) -> float:
        import torch
        import torch.distributed as dist
        if self.world == 1:
            return x
        t = torch.tensor([x], device=self.dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())


def committed_traffic(variant: str, frames: int):
    """DRAM bytes per step from the committed ncu capture of this workload (profiles/), if one matches."""
    for name in ("r2_dram_traffic.json", "r1_tc3_dram_traffic.json"):
        p = os.path.join(ROOT, "profiles", name)
        if not os.path.exists(p):
            continue
        tj = json.load(open(p))
        if tj.get("variant", "lstm") == variant and tj.get("frames") == frames:
            tot = sum(u["dram_read_bytes"] + u["dram_write_bytes"] for u in tj["units"].values())
            return tot, f"profiles/{name} (ncu --set full, same command line, not measured in this run)", tj["units"]
    return None, None, {}


def offline_record(ctx: Ctx, blob: bytes, variant: str, B: int, n_samples: int, steps: int, warmup: int, profile: bool):
    import torch
    from nunet_b200.engine import NunetEngine
    from nunet_b200.synth import synth_clips
    T = 1 + (n_samples - 512) // 256
    dev = ctx.dev
    eng = NunetEngine(blob, max_frames=B * T, device=ctx.local_rank, ctfa_mode="causal_avg32", variant=1 if variant == "ddb" else 0)
    pool = synth_clips(min(B, 32), n_samples, first_clip=1000 * ctx.rank)
    wav_h = torch.from_numpy(np.tile(pool, ((B + len(pool) - 1) // len(pool), 1))[:B]).contiguous().pin_memory()
    wav_d = wav_h.to(dev)
    out_d = torch.empty((B, (T - 1) * 256 + 512), device=dev, dtype=torch.float32)
    out_h = torch.empty(out_d.shape, dtype=torch.float32).pin_memory()
    W = max(warmup, 3)
    for _ in range(W):
        eng.forward_wav_into(wav_d, out_d)
    launches_per_step = eng.last_launch_count
    sampler = ClockSampler(ctx.local_rank)
    ctx.barrier()
    if ctx.rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        eng.forward_wav_into(wav_d, out_d)
    e1.record()
    ctx.barrier()
    ms = ctx.max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if ctx.rank == 0 else None
    frames_total = ctx.world * B * T * steps
    value = frames_total / (ms * 1e-3)

    # end to end through the host entry point (H2D + kernels + D2H per step)
    for _ in range(2):
        eng.forward_wav_host(wav_h, out_h)
    ctx.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        eng.forward_wav_host(wav_h, out_h)
    torch.cuda.synchronize(dev)
    e2e_value = frames_total / ctx.max_over_ranks(time.perf_counter() - t0)

    roofline = None
    if ctx.rank == 0:
        peak, peak_src = measured_peaks()
        b_alg = B_ALG[variant]
        path_gbs = value / ctx.world * b_alg / 1e9
        traffic, traffic_src, units = committed_traffic(variant, B * T)
        roofline = {"bound": "hbm", "kernel": "whole path (all launches of one step)", "achieved": path_gbs, "peak": peak,
                    "unit": "GB/s", "frac": path_gbs / peak, "traffic": traffic, "traffic_source": traffic_src,
                    "peak_source": peak_src, "b_alg_bytes_per_frame": b_alg, "algorithmic_bytes": b_alg * B * T,
                    "fp32_equiv_tflops": value / ctx.world * FLOP_PER_FRAME[variant] / 1e12}
        # second roofline of the path: fp32-grade results cost three fp16 tensor products per multiply-add
        tpeak, tsrc = measured_tensor_peak()
        t3 = 3.0 * value / ctx.world * FLOP_PER_FRAME[variant] / 1e12
        roofline["tensor_3product"] = {"bound": "tensor", "achieved": t3, "peak": tpeak, "unit": "TFLOP/s", "frac": t3 / tpeak, "peak_source": tsrc,
                                       "note": "issued fp16 tensor FLOP/s = 3 x the model's FLOPs (split-half hi/lo products)"}
        if profile:
            # per-launch CUDA events (one extra step) -> the launch with the largest share of the step
            eng.profile(True)
            eng.forward_wav_into(wav_d, out_d)
            torch.cuda.synchronize(dev)
            ent = eng.profile_entries()
            eng.profile(False)
            total_ms = sum(e[1] for e in ent)
            top = max(ent, key=lambda e: e[1])
            u = units.get(top[0])
            roofline["dominant_kernel"] = {
                "kernel": top[0], "kernel_ms": top[1], "kernel_share_of_step": top[1] / total_ms,
                "algorithmic_bytes": top[2], "achieved": top[2] / (top[1] * 1e-3) / 1e9,
                "frac": top[2] / (top[1] * 1e-3) / 1e9 / peak,
                "traffic": (u["dram_read_bytes"] + u["dram_write_bytes"]) if u else None}
            by_kernel = {}
            for name, kms, nbytes in ent:
                k = by_kernel.setdefault(name.split(":")[-1], [0.0, 0.0, 0])
                k[0] += kms
                k[1] += nbytes
                k[2] += 1
            roofline["by_kernel_ms"] = {n: round(v[0], 3) for n, v in sorted(by_kernel.items(), key=lambda kv: -kv[1][0])}
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            with open(os.path.join(ROOT, "gpurun_out", f"bench_kernel_profile_{variant}.json"), "w") as f:
                json.dump({"total_ms": total_ms, "entries": ent}, f)
    eng.close()
    del eng, wav_d, out_d
    torch.cuda.empty_cache()
    return {"value": value, "ms_per_step": ms / steps, "steps": steps, "warmup": W, "clocks": clocks,
            "rtf_per_stream_equivalent": (1.0 / (value / ctx.world)) / 0.016,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(wav_h.numel() * 4),
                    "d2h_bytes_per_step": int(out_h.numel() * 4)},
            "gpu_launches": int(launches_per_step * steps), "launches_per_step": int(launches_per_step), "roofline": roofline}


def streaming_record(ctx: Ctx, blob: bytes, variant: str, S: int, steps: int, warmup: int):
    """configs[2]: S concurrent streams, frame-basis, history carried on the device (interpreter_proposed.py:200-366 for S
    streams at once).  One step = one hop of every stream."""
    import torch
    from nunet_b200.engine import NunetEngine
    from nunet_b200.synth import synth_clips
    dev = ctx.dev
    eng = NunetEngine(blob, max_streams=S, device=ctx.local_rank, ctfa_mode="frame_div32", dc_mode="edge",
                      variant={"lstm": 0, "ddb": 1, "hybrid": 2}[variant])
    variant = "lstm" if variant == "hybrid" else variant       # same graph, same algorithmic bytes
    W = max(warmup, 4)       # 2 plain steps + one CUDA-graph capture per step parity happen before the timed region
    nh = min(W + steps + 4, 260)
    pool = synth_clips(min(S, 32), 256 * nh, first_clip=1000 * ctx.rank)
    hops_h = torch.from_numpy(np.tile(pool, ((S + len(pool) - 1) // len(pool), 1))[:S]).reshape(S, nh, 256)
    hops_h = hops_h.permute(1, 0, 2).contiguous().pin_memory()          # [hop][stream][256]
    hops_d = hops_h.to(dev)
    out_d = torch.empty((S, 256), device=dev)
    out_h = torch.empty((S, 256)).pin_memory()
    eng.stream_reset()
    for i in range(W):
        eng.stream_step_wav(hops_d[i % nh], out_d)
    launches = eng.last_launch_count
    sampler = ClockSampler(ctx.local_rank)
    ctx.barrier()
    if ctx.rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        eng.stream_step_wav(hops_d[(W + i) % nh], out_d)
    e1.record()
    ctx.barrier()
    ms = ctx.max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if ctx.rank == 0 else None
    value = ctx.world * S * steps / (ms * 1e-3)
    ctx.barrier()
    t0 = time.perf_counter()
    for i in range(steps):
        eng.stream_step_wav_host(hops_h[(W + i) % nh], out_h)
    torch.cuda.synchronize(dev)
    e2e = ctx.world * S * steps / ctx.max_over_ranks(time.perf_counter() - t0)
    peak, peak_src = measured_peaks()
    b_alg = B_ALG_STREAM[variant]
    gbs = value / ctx.world * b_alg / 1e9
    eng.close()
    del eng
    torch.cuda.empty_cache()
    return {"value": value, "ms_per_step": ms / steps, "steps": steps, "warmup": W, "clocks": clocks,
            "rtf_per_stream": (ms / steps * 1e-3) / 0.016,
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": S * 256 * 4, "d2h_bytes_per_step": S * 256 * 4},
            "gpu_launches": int(launches * steps), "launches_per_step": int(launches),
            "roofline": {"bound": "hbm", "kernel": "whole streaming step (one CUDA-graph replay)", "achieved": gbs, "peak": peak,
                         "unit": "GB/s", "frac": gbs / peak, "traffic": None, "traffic_source": None, "peak_source": peak_src,
                         "b_alg_bytes_per_frame": b_alg, "algorithmic_bytes": b_alg * S}}


def offline_config(variant, B, seconds, T):
    name = "configs[1]" if variant == "lstm" else "configs[3]"
    model = "NUNet-TLS-LSTM" if variant == "lstm" else "NUNet-TLS dilated-dense bottleneck variant"
    return {"workload": f"{name}: offline batch={B} x {seconds:g} s @16 kHz clips per GPU, {model}, T={T} frames/clip, "
                        "ctfa causal_avg32",
            "clips_per_gpu": B, "frames_per_clip": T,
            "weights": "trained nutls_lstm.h5 (reference checkpoint)" if variant == "lstm"
            else "shipped nutls.tflite, int8 tensors dequantised (the variant's only checkpoint)",
            "l2_policy": "working set (inputs 65 MB + >30 GB activations per pass) far exceeds the 126 MB L2"}


def streaming_config(variant, S):
    return {"workload": f"configs[2]: streaming frame-basis, {S} concurrent streams per GPU with carried conv / "
                        f"{'LSTM' if variant == 'lstm' else 'DDB'} history, NUNet-TLS{'-LSTM' if variant == 'lstm' else ''}, "
                        "ctfa frame_div32 (one-frame graph)",
            "streams_per_gpu": S,
            "weights": "trained nutls_lstm.h5 (reference checkpoint)" if variant == "lstm" else "shipped nutls.tflite (dequantised)",
            "l2_policy": f"per-step working set {S * B_ALG_STREAM[variant] / 1e9:.1f} GB exceeds the 126 MB L2"}


def load_weights_or_die(variant):
    """Headline numbers never run on random weights: a missing blob is an error."""
    from nunet_b200.weights import load_ddb_weights, load_default_weights
    try:
        return load_ddb_weights() if variant == "ddb" else load_default_weights()
    except FileNotFoundError as e:
        print(f"bench.py: weight blob for variant '{variant}' is missing ({e}); refusing to run on random weights", file=sys.stderr)
        raise SystemExit(3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="clips per GPU per step")
    ap.add_argument("--seconds", type=float, default=4.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--config", default="all", choices=["all", "offline", "streaming"],
                    help="all (default) = headline configs[1] + extra records for configs[2] and configs[3] (extras on 1 GPU "
                         "only); offline = configs[1] alone (configs[3] with --variant ddb); streaming = configs[2] alone")
    ap.add_argument("--streams", type=int, default=1024)
    ap.add_argument("--variant", default="lstm", choices=["lstm", "ddb"])
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    n_samples = int(round(args.seconds * 16000))
    T = 1 + (n_samples - 512) // 256
    variant = args.variant
    from nunet_b200.weights import pack_blob
    weights = load_weights_or_die(variant)

    base = {"metric": METRIC, "unit": UNIT, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": DTYPE,
            "dtype_note": DTYPE_NOTE, "data": "synthetic"}

    # ------------------------------------------------------------------ reference arm: the CPU oracle on host cores
    if args.impl == "reference":
        if rank != 0:
            return
        sample_clips = 2
        steps, warmup = max(1, args.steps), max(0, args.warmup)
        val, sec, cores, sample = cpu_offline_leg(weights, n_samples, sample_clips, steps, warmup, variant=variant)
        cfg = offline_config(variant, args.batch, args.seconds, T)
        cfg["reference_sample_clips_per_step"] = sample_clips     # frames/s is size-normalised; the CPU arm steps over 2 clips
        rec = dict(base, impl="reference", value=val, n_gpus=args.gpus, steps=steps, warmup=warmup, ms_per_step=sec * 1e3, config=cfg,
                   cpu_baseline={"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
                   e2e={"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                   note="TensorFlow/TFLite cannot be installed offline; this is the CPU restatement of the reference path "
                        f"(oracle/) on ONE host ({cores} threads) whatever --gpus says; each step = {sample_clips} clips x {T} frames")
        if not args.no_cpu_baseline:
            try:
                rec["cpu_baseline"]["streaming_form"] = {
                    "what": "oracle frame loop, batch 1, one frame per call (interpreter_proposed.py:200-366)",
                    "k1": cpu_streaming_leg(weights, 40, 1, variant), "kall": cpu_streaming_leg(weights, 40, cores, variant)}
            except Exception as e:
                rec["cpu_baseline"]["streaming_form"] = {"error": repr(e)}
        print(json.dumps(rec))
        return

    # ------------------------------------------------------------------ our arm
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    ctx = Ctx(rank, local_rank, world)
    blob = pack_blob(weights, 1 if variant == "ddb" else 0)
    if world > 1:
        from nunet_b200.sharding import broadcast_blob
        dist.init_process_group("nccl", device_id=ctx.dev)
        # the only collective of the path: rank 0's packed weights go to every GPU once (NVLink / NCCL)
        blob = broadcast_blob(blob if rank == 0 else None, src=0, device=ctx.dev)

    if args.config == "streaming":
        r = streaming_record(ctx, blob, variant, args.streams, args.steps, args.warmup)
        line = dict(base, n_gpus=world, config=streaming_config(variant, args.streams), cpu_baseline=None, **r)
    else:
        r = offline_record(ctx, blob, variant, args.batch, n_samples, args.steps, args.warmup, profile=True)
        line = dict(base, n_gpus=world, config=offline_config(variant, args.batch, args.seconds, T), **r)
        if args.config == "all" and world == 1 and variant == "lstm":
            extra = {}
            try:
                s_steps = max(args.steps, 100)           # a streaming step is ~2 ms: time at least 100 of them
                sr = streaming_record(ctx, blob, "lstm", args.streams, s_steps, args.warmup)
                extra["configs[2]"] = dict(config=streaming_config("lstm", args.streams), metric=METRIC, unit=UNIT, **sr)
            except Exception as e:
                extra["configs[2]"] = {"error": repr(e)}
            try:
                # not a BASELINE config: the same streaming workload with the arithmetic of the DEPLOYED int8 dynamic-range graph
                # (engine variant 2; FP32 / dp4a SIMT kernels, every hybrid operator's input quantised per stream and call)
                from nunet_b200.tflite_export import hybrid_weight_set
                hs = min(args.streams, 256)
                hr = streaming_record(ctx, pack_blob(hybrid_weight_set(weights), 2), "hybrid", hs, max(args.steps, 20), args.warmup)
                cfg_h = streaming_config("lstm", hs)
                cfg_h["workload"] = cfg_h["workload"].replace("configs[2]", "deployed int8-hybrid arithmetic (SURVEY 8(f)3)")
                extra["int8_hybrid_streaming"] = dict(config=cfg_h, metric=METRIC, unit=UNIT, **hr)
            except Exception as e:
                extra["int8_hybrid_streaming"] = {"error": repr(e)}
            try:
                dw = load_weights_or_die("ddb")
                dr = offline_record(ctx, pack_blob(dw, 1), "ddb", args.batch, n_samples, args.steps, args.warmup, profile=True)
                extra["configs[3]"] = dict(config=offline_config("ddb", args.batch, args.seconds, T), metric=METRIC, unit=UNIT, **dr)
            except Exception as e:
                extra["configs[3]"] = {"error": repr(e)}
            line["extra"] = extra
        if rank == 0 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_record(weights, n_samples, variant)
        else:
            line["cpu_baseline"] = None
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
