#!/usr/bin/env python
"""bench.py -- jivetalking four-pass chain (analyse / process / measure / normalise) on B200.

Metric (BASELINE.json): samples/s (x realtime) through the full 4-pass chain.  One "step" = the
whole chain over one 60 min 48 kHz mono f32 stream (BASELINE.json configs[1]) per GPU; with N GPUs
each rank processes its own file (configs[2], file-per-GPU, no data-path collective -> weak
scaling).  `value` is timed with the stream already resident in HBM (jt_process_audio_dev); `e2e`
goes through the reference-facing C ABI call jt_process_audio with pinned HOST buffers, so the
host->device copy of the PCM and the device->host copy of the 16-bit result are inside the timed
region.  `--impl reference` times the CPU oracle (the restated FFmpeg path; the reference's own
embedded-FFmpeg path needs Go + libffmpeg.a, absent here) on the box's host cores.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RATE = 48000
MINUTES = 60
BYTES_PER_SAMPLE_4PASS = 15.35          # 4 + 5.8375 + 1.8375 + 3.675
# algorithmic HBM bytes per input sample of each kernel group (bytes of its resident input + output)
KERNEL_BYTES = {
    "anlmdn": 8.0,                      # f32 in + f32 out
    "afftdn": 8.0,
    "adeclick": 16.0 * 0.91875,         # f64 in + f64 out at 44.1 kHz
    "truepeak_oversample": 4.0,
    "swr_resample": 8.0 + 8.0 * 0.91875,
    "astats": 8.0,
    "aspectralstats": 4.0,
    "r128_kweight_ticks": 4.0,
    "envelope_follower": 16.0,
    "agate_gain": 24.0, "acompressor_gain": 24.0,
    "alimiter": 16.0 * 0.91875,
    "biquad": 8.0, "convert": 12.0, "raw_frame_stats": 4.0, "deesser": 16.0,
    "loudnorm_linear_gain": 16.0 * 0.91875, "volume": 8.0,
}


# dram__bytes_read.sum + dram__bytes_write.sum per launch on the 60 min stream, from the `ncu --set full` captures
# summarised in profiles/ncu_r1k.md / ncu_r1h.md (cold-cache replay; None where the kernel was not captured)
NCU_TRAFFIC_BYTES_60MIN = {
    "adeclick:interp": 1.661498e9 + 0.722141e9,            # profiles/ncu_r1k.md
    "anlmdn": 0.706247e9 + 0.651037e9,                     # profiles/ncu_r1k.md
    "afftdn:fwd": 0.702286e9 + 2.324030e9,
    "alimiter": 3.293223e9 + 1.231703e9,
    "swr_resample:qlane_f64": 0.705825e9 + 1.226726e9,
    "envelope_follower": 8.088282e9 + 1.336265e9,          # profiles/ncu_r1k.md
    "r128_kweight_ticks": 1.383083e9 + 0.006421e9,         # profiles/ncu_r1d.md
}
# what actually bounds each group; the contract's roofline is reported against HBM regardless
KERNEL_BOUND = {
    "adeclick:interp": "latency (10 warps/SM, shared-memory round trips of the banded LDL^T k-loop)",
    "anlmdn": "FP32 issue / shared-memory bandwidth",
    "envelope_follower": "f64 dependent-issue latency (one warp per scheduler)",
    "afftdn:fwd": "barriers of the shared-memory FFT + f64 tracking statistics",
}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.rows, self.stop = device, [], threading.Event()
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.t.join(timeout=6)

    def summary(self):
        sm = sorted(int(float(r[1])) for r in self.rows if len(r) > 2 and r[1].replace(".", "").isdigit())
        mx = [int(float(r[2])) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def make_input(seed, minutes, kind="speech"):
    from jivetalking_b200 import synth
    import numpy as np
    gen = synth.speech_like if kind == "speech" else synth.podcast_like      # podcast: > 20 % room tone, elects both regions
    blocks = [gen(600.0 if m + 10 <= minutes else (minutes - m) * 60.0, RATE, seed=seed * 1000 + m)
              for m in range(0, minutes, 10)]
    return np.concatenate(blocks)


def cpu_chain(x, rate):
    """The oracle's four-pass chain (test infrastructure; only the cpu_baseline / reference legs call it)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_graph as OG
    from jivetalking_b200 import gpudsp
    OG.run_spec(gpudsp.pass1_spec(), x, rate, want_pcm=False)
    p2 = OG.run_spec(gpudsp.default_pass2_spec(), x, rate)
    last = [m for m in p2["meta"] if not math.isnan(m["I"])][-1]
    out_tp = -120.0 if last["true_peak"] <= 0 else 20 * math.log10(last["true_peak"])
    spec3, plan = gpudsp.build_pass3_spec(last["I"], out_tp)
    p3 = OG.run_spec(spec3, p2["pcm"], 44100, want_pcm=False)
    st = gpudsp.LoudnormStats()
    for k in ("input_i", "input_tp", "input_lra", "input_thresh"):
        setattr(st, k, p3["loudnorm"][k])
    spec4, _, _ = gpudsp.build_pass4_spec(plan, st)
    return OG.run_spec(spec4, p2["pcm"], 44100)["pcm"]


def _cpu_worker(args):
    seed, seconds = args
    from jivetalking_b200 import synth
    x = synth.speech_like(seconds, RATE, seed=seed)
    t0 = time.perf_counter()
    cpu_chain(x, RATE)
    return len(x), time.perf_counter() - t0


def cpu_baseline(cores, seconds_per_core):
    """Each host core runs the scalar oracle chain over its own `seconds_per_core` s stream
    (the reference's own parallelism is one file per CPU thread: cmd/jivetalking/pool.go:122-153)."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, [(900 + i, seconds_per_core) for i in range(cores)])
    wall = time.perf_counter() - t0
    total = sum(r[0] for r in res)
    busy = max(r[1] for r in res)
    return total / busy, total, busy, wall


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sec = 150.0                           # ~20 s of scalar CPU work per core (the oracle chain runs ~8x realtime/core)
    vals = []
    for i in range(args.warmup + args.steps):
        v, total, busy, _ = cpu_baseline(cores, sec)
        if i >= args.warmup:
            vals.append((v, busy))
    value = sum(v for v, _ in vals) / len(vals)
    line = {"impl": "reference", "metric": "samples/s full 4-pass chain", "value": value, "unit": "samples/s",
            "realtime_x": value / RATE, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * sum(b for _, b in vals) / len(vals), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32/f64", "data": "synthetic",
            "config": {"workload": "60 min 48 kHz mono f32, full 4-pass chain (BASELINE.json configs[1])",
                       "note": "bounded sample of that workload per step"},
            "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": "port",
                             "sample": f"{cores} x {sec:.0f} s streams of the C2 recipe, one scalar oracle chain per core"},
            "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--minutes", type=int, default=MINUTES)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-adaptive", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    from jivetalking_b200 import gpudsp

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the chain has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from jivetalking_b200 import shard
    my_files = shard.assign_files(world, rank, world)               # one file per GPU (configs[2])
    x = make_input(shard.stream_seed(12345, my_files[0]), args.minutes)   # C2 / C3: seeds 12345 + file index
    n = len(x)
    h_in = torch.from_numpy(x).pin_memory()
    d_in = h_in.cuda(non_blocking=False)
    out_cap = int(n * 44100 / RATE) + 3 * 4096
    d_out = torch.empty(out_cap, dtype=torch.int16, device="cuda")
    h_out = torch.empty(out_cap, dtype=torch.int16).pin_memory()
    ctx = gpudsp.Context(local)
    lib_stream = torch.cuda.ExternalStream(ctx.cuda_stream(), device=torch.device("cuda", local))   # the stream the kernels run on

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_dev():
        return ctx.process_audio_ptr(d_in.data_ptr(), n, RATE, 1, gpudsp.FMT_FLT, d_out.data_ptr(), out_cap, True)

    def step_e2e():
        return ctx.process_audio_ptr(h_in.data_ptr(), n, RATE, 1, gpudsp.FMT_FLT, h_out.data_ptr(), out_cap, False)

    # ---- device-resident timing (value) -------------------------------------------------------
    ctx.enable_timing(True)                  # warm up in the configuration that is timed (event pools, caches)
    for _ in range(args.warmup):
        res = step_dev()
    ctx.reset_counters()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        t0 = time.perf_counter()
        ev0.record(lib_stream)
        for _ in range(args.steps):
            res = step_dev()
        ev1.record(lib_stream)
        torch.cuda.synchronize()
        barrier()
        t_wall = time.perf_counter() - t0
    # device clock (CUDA events on the library's own stream) around the K steps, host gaps between kernels included
    t_dev = ev0.elapsed_time(ev1) * 1e-3
    launches = ctx.launch_count()
    timings = ctx.kernel_timings()
    ctx.enable_timing(False)

    # ---- end to end through the C ABI with host buffers (e2e) ----------------------------------
    for _ in range(max(1, args.warmup // 2)):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    ev0.record(lib_stream)
    for _ in range(args.steps):
        res_e = step_e2e()
    ev1.record(lib_stream)
    barrier()
    t_e2e_wall = time.perf_counter() - t0
    t_e2e = max(ev0.elapsed_time(ev1) * 1e-3, 0.0)

    # ---- the adaptive path on the library side (Pass 1 -> detector -> 17 band graphs -> AdaptConfig -> Pass 2..4), reported
    #      next to the headline; Pass 2 cannot overlap Pass 1 here because its spec depends on Pass 1's measurements ----------
    adaptive = None
    if rank == 0 and not args.no_adaptive:
        from jivetalking_b200 import adapt
        # a conversational recipe (speech runs / room-tone pauses) so the detector elects both regions and every adaptive branch runs
        # (one 10 min block tiled: generating it costs host seconds, not GPU work)
        blk = make_input(4242, min(10, args.minutes), kind="podcast")
        d_in.copy_(torch.from_numpy(np.tile(blk, (n + len(blk) - 1) // len(blk))[:n]))
        for _ in range(2):
            adapt.process_audio_adaptive_ptr(ctx, d_in.data_ptr(), n, RATE, 1, gpudsp.FMT_FLT, d_out.data_ptr(), out_cap, True)
        torch.cuda.synchronize()
        ev0.record(lib_stream)
        for _ in range(args.steps):
            res_a, an_a = adapt.process_audio_adaptive_ptr(ctx, d_in.data_ptr(), n, RATE, 1, gpudsp.FMT_FLT, d_out.data_ptr(), out_cap, True)
        ev1.record(lib_stream)
        torch.cuda.synchronize()
        t_ad = ev0.elapsed_time(ev1) * 1e-3 / args.steps
        va = an_a.voice_activity
        adaptive = {"value": n / t_ad, "unit": "samples/s", "realtime_x": n / t_ad / RATE, "ms_per_step": 1e3 * t_ad,
                    "entry": "jt_process_audio_adaptive_dev (ProcessAudio with AnalyseAudio + AdaptConfig + MeasureOutputRegions inside the library)",
                    "input": f"{args.minutes} min conversational synthetic (synth.podcast_like, 10 min block tiled), device-resident",
                    "regions_remeasured": int(an_a.filtered_regions.has_room_tone + an_a.filtered_regions.has_speech +
                                              an_a.final_regions.has_room_tone + an_a.final_regions.has_speech),
                    "pass2_spec": an_a.pass2_spec.decode(), "speech_profile": bool(va.has_speech_profile),
                    "noise_profile": bool(va.has_noise_profile), "voice_activated": bool(va.voice_activated),
                    "noise_floor": va.floor, "final_lufs": res_a.final.input_i, "final_dbtp": res_a.final.input_tp}
    # ---- the result's container: FLAC stream of the s16 44.1 kHz output, device resident -------------------
    flac = None
    if rank == 0 and not args.no_adaptive:
        res_f = step_dev()
        n_out = int(res_f.n_out)
        cap_b = int(gpudsp.lib().jt_flac_max_bytes(n_out, 4096))
        d_flac = torch.empty(cap_b, dtype=torch.uint8, device="cuda")
        ctx.flac_encode_ptr(d_out.data_ptr(), n_out, 44100, 4096, d_flac.data_ptr(), cap_b, True)
        torch.cuda.synchronize()
        ev0.record(lib_stream)
        for _ in range(args.steps):
            nbytes = ctx.flac_encode_ptr(d_out.data_ptr(), n_out, 44100, 4096, d_flac.data_ptr(), cap_b, True)
        ev1.record(lib_stream)
        torch.cuda.synchronize()
        t_fl = ev0.elapsed_time(ev1) * 1e-3 / args.steps
        flac = {"ms_per_step": 1e3 * t_fl, "realtime_x": (n_out / 44100) / t_fl, "samples_per_s": n_out / t_fl,
                "bytes": int(nbytes), "ratio": nbytes / (2.0 * n_out),
                "algorithmic_GBps": (2.0 * n_out + nbytes) / t_fl / 1e9,
                "entry": "jt_flac_encode_dev (4096-sample frames, fixed predictors + partitioned Rice; encoder.go:92-101)"}
        del d_flac
    barrier()

    if world > 1:
        tt = torch.tensor([t_dev, t_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_dev, t_e2e = float(tt[0]), float(tt[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    total_samples = n * world * args.steps
    value = total_samples / t_dev
    e2e = total_samples / t_e2e
    peak, peak_src = measured_peaks()
    timings.sort(key=lambda t: -t[1])
    gaps = [t for t in timings if t[0].startswith("gap:")]
    timings = [t for t in timings if not t[0].startswith("gap:")]
    gpu_timings = [t for t in timings if not t[0].startswith("host:")]
    dom = gpu_timings[0] if gpu_timings else ("none", 0.0, 0)
    dom_ms_per_step = dom[1] / args.steps
    dom_bytes = KERNEL_BYTES.get(dom[0].split(":")[0], 8.0) * n
    achieved = dom_bytes / (dom_ms_per_step * 1e-3) / 1e9 if dom_ms_per_step > 0 else 0.0
    kernel_ms = sum(t[1] for t in gpu_timings) / args.steps
    line = {
        "metric": "samples/s full 4-pass chain", "value": value, "unit": "samples/s", "realtime_x": value / RATE,
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_dev / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32/f64", "data": "synthetic",
        "config": {"workload": f"{args.minutes} min 48 kHz mono f32 per GPU, full 4-pass chain (BASELINE.json configs[1]; file per GPU = configs[2])",
                   "samples_per_gpu": n, "pass2_spec": "DefaultFilterConfig (filters.go:353-355)",
                   "l2": "inputs (691 MB/stream) and every intermediate exceed the 126 MB L2"},
        "e2e": {"value": e2e, "unit": "samples/s", "realtime_x": e2e / RATE, "h2d_bytes_per_step": int(n * 4),
                "d2h_bytes_per_step": int(res_e.n_out * 2), "ms_per_step": 1e3 * t_e2e / args.steps},
        "gpu_launches": int(launches),
        "timing": {"method": "CUDA events on the library stream around the K steps (max over ranks)",
                   "wall_ms_per_step": 1e3 * t_wall / args.steps, "e2e_wall_ms_per_step": 1e3 * t_e2e_wall / args.steps},
        "roofline": {"bound": "hbm", "kernel": dom[0], "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak if peak else None,
                     "traffic": NCU_TRAFFIC_BYTES_60MIN.get(dom[0]) if args.minutes == 60 else None,
                     "traffic_source": "profiles/ncu_r1k.md / ncu_r1h.md (bytes per launch, ncu --set full)", "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": dom_bytes,
                     "kernel_ms_per_step": dom_ms_per_step, "kernel_share_of_step": dom_ms_per_step / (1e3 * t_dev / args.steps),
                     "chain_frac": (BYTES_PER_SAMPLE_4PASS * n / (t_dev / args.steps) / 1e9) / peak,
                     "note": "not HBM bound: " + KERNEL_BOUND.get(dom[0], "latency / issue bound (DESIGN.md section 5)")},
        "kernels_ms_per_step": {t[0]: round(t[1] / args.steps, 3) for t in timings},
        "kernel_ms_total_per_step": kernel_ms,
        "device_idle_ms_per_step": {g[0][4:]: round(g[1] / args.steps, 3) for g in sorted(gaps, key=lambda g: -g[1])[:12]},
        "device_idle_ms_total_per_step": sum(g[1] for g in gaps) / args.steps,
        "clocks": clocks.summary(),
        "result": {"final_lufs": res.final.input_i, "final_dbtp": res.final.input_tp, "final_lra": res.final.input_lra,
                   "n_out": int(res.n_out), "limiter_needed": int(res.limiter_needed), "pass4_type": int(res.pass4.normalization_type)},
    }
    if adaptive is not None:
        line["adaptive"] = adaptive
    if flac is not None:
        line["flac_encode"] = flac
    if not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        sec = 150.0                           # ~20 s of scalar CPU work per core (the oracle chain runs ~8x realtime/core)
        v, total, busy, wall = cpu_baseline(cores, sec)
        line["cpu_baseline"] = {"value": v, "unit": "samples/s", "realtime_x": v / RATE, "cores": cores, "kind": "port",
                                "sample": f"{cores} x {sec:.0f} s streams of the same recipe, one scalar oracle chain per core "
                                          f"({busy:.1f} s of CPU work per core)"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
