#!/usr/bin/env python
"""bench.py -- jivetalking ProcessAudio (analyse / process / measure / normalise) on B200.

Metric (BASELINE.json): samples/s (x realtime) through the full 4-pass chain.  One "step" = ProcessAudio
(processor.go:78-216: Pass 1 -> detector -> 17 band graphs -> AdaptConfig -> Pass 2 -> region re-measures -> Pass 3 ->
Pass 4 -> region re-measures) over one 60 min 48 kHz mono f32 conversational stream (BASELINE.json configs[1]) per GPU; with
N GPUs each rank processes its own file (configs[2], file per GPU, no data-path collective -> weak scaling).

  value : the stream already resident in HBM (jt_process_audio_adaptive_dev), CUDA events on the library's stream
  e2e   : the reference-facing C-ABI call jt_process_audio_adaptive with pinned HOST buffers -- the host->device copy of the
          PCM and the device->host copy of the 16-bit result are inside the timed region
  extras: fixed_spec (caller-supplied DefaultFilterConfig spec, the round-1 headline), flac_encode (the result's container),
          batch_sweep (configs[4]: 1..256 concurrent 10 min streams, N = 1 only), stream_sharded (configs[3]: one 3 h
          96 kHz stereo stream over the N GPUs, N > 1 only)

`--impl reference` times the CPU oracle's ProcessAudio (oracle/chain_oracle.py: the restated FFmpeg path -- the reference's
own embedded-FFmpeg path needs Go + libffmpeg.a, absent here) on the box's host cores; it never loads libjtdsp.
"""
import argparse
import importlib.util
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))

RATE = 48000
MINUTES = 60
BYTES_PER_SAMPLE_4PASS = 15.35          # 4 (Pass 1) + 5.8375 (Pass 2) + 1.8375 (Pass 3) + 3.675 (Pass 4)
# algorithmic HBM bytes per input sample of each kernel group: bytes of its resident input + output
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
    "loudnorm_linear_gain": 16.0 * 0.91875, "volume": 8.0, "band_rms": 4.0,
}
KERNEL_BOUND = {
    "adeclick:interp": "latency (banded LDL^T k-loop, shared-memory round trips)",
    "anlmdn": "FP32 issue / shared-memory bandwidth",
    "anlmdn:screen": "FP32 issue (a subtract and an FMA per sample and lag; the exact kernel walks only the listed hops)",
    "envelope_follower": "f64 dependent-issue latency (one lane per stream segment)",
    "afftdn:fwd": "barriers of the shared-memory FFT + f64 tracking statistics",
}


def load_by_path(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def synth_module():
    """the synthetic-input generators, loaded by path: the reference arm must not import the product package"""
    return load_by_path("jt_synth", os.path.join(ROOT, "jivetalking_b200", "synth.py"))


def ncu_traffic():
    """dram bytes per launch of each kernel on the 60 min stream, from the committed `ncu --set full` summaries
    (profiles/kernel_roofline_r2.json, written by scripts/summarise_profiles.py)"""
    p = os.path.join(ROOT, "profiles", "kernel_roofline_r2.json")
    try:
        return {k: v.get("dram_bytes_per_launch") for k, v in json.load(open(p)).get("kernels", {}).items()}, os.path.relpath(p, ROOT)
    except Exception:
        return {}, None


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


def _block(args):
    seed, seconds = args
    return synth_module().podcast_like(seconds, RATE, seed=seed)


def make_input(seed, minutes):
    """C2: `minutes` of the conversational recipe (speech runs, room-tone pauses: the detector elects both regions and every
    adaptive branch runs), 10 min blocks with distinct seeds, generated on a few host processes"""
    import multiprocessing as mp
    import numpy as np
    jobs = [(seed * 1000 + m, 600.0 if m + 10 <= minutes else (minutes - m) * 60.0) for m in range(0, minutes, 10)]
    with mp.get_context("spawn").Pool(min(len(jobs), max(1, (os.cpu_count() or 2) // 2))) as pool:
        return np.concatenate(pool.map(_block, jobs))


# ---- CPU legs: the oracle's ProcessAudio on host cores (test infrastructure; never the thing shipped) --------------------
def _cpu_worker(args):
    seed, seconds = args
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import chain_oracle as CO
    x = synth_module().podcast_like(seconds, RATE, seed=seed)
    t0 = time.perf_counter()
    CO.process_audio(x, RATE)
    return len(x), time.perf_counter() - t0


def cpu_baseline(cores, seconds_per_core):
    """Each host core runs the scalar oracle ProcessAudio over its own stream (the reference's own parallelism is one file per
    CPU thread: cmd/jivetalking/pool.go:122-153)."""
    import multiprocessing as mp
    with mp.get_context("spawn").Pool(cores) as pool:
        res = pool.map(_cpu_worker, [(900 + i, seconds_per_core) for i in range(cores)])
    total = sum(r[0] for r in res)
    busy = max(r[1] for r in res)
    return total / busy, total, busy


CPU_SAMPLE_SECONDS = 120.0              # ~25 s of scalar CPU work per core per step


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    vals = []
    for i in range(args.warmup + args.steps):
        v, total, busy = cpu_baseline(cores, CPU_SAMPLE_SECONDS)
        if i >= args.warmup:
            vals.append((v, busy))
    value = sum(v for v, _ in vals) / len(vals)
    line = {"impl": "reference", "metric": "samples/s full 4-pass chain", "value": value, "unit": "samples/s",
            "realtime_x": value / RATE, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * sum(b for _, b in vals) / len(vals), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32/f64", "data": "synthetic",
            "config": {"workload": "ProcessAudio (adaptive), 60 min 48 kHz mono f32 (BASELINE.json configs[1])",
                       "note": "bounded sample of that workload per step"},
            "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": "port",
                             "sample": f"{cores} x {CPU_SAMPLE_SECONDS:.0f} s conversational streams, one scalar oracle ProcessAudio per core "
                                       "(Pass 1, detector, 17 bands, AdaptConfig, Pass 2-4, region re-measures)"},
            "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ---- configs[4]: batch sweep ------------------------------------------------------------------------------------------
def batch_sweep(local, torch, gpudsp, adapt, steps_cap=256):
    """B concurrent 10 min streams on ONE GPU, B = 1 .. 256: a pool of worker threads, one jt_ctx (one CUDA stream) each -- the
    reference's goroutine-per-file pool (pool.go:122-153) -- drains the B files.  x realtime = B x 600 s / wall time, device
    synchronised on both sides.  The B inputs are four distinct seeds used round-robin, device resident."""
    import numpy as np
    syn = synth_module()
    n = 600 * RATE
    inputs = [torch.from_numpy(syn.podcast_like(600.0, RATE, seed=777 + i)).cuda() for i in range(4)]
    out_cap = int(n * 44100 / RATE) + 3 * 4096
    max_workers = 8
    ctxs = [gpudsp.Context(local) for _ in range(max_workers)]
    outs = [torch.empty(out_cap, dtype=torch.int16, device="cuda") for _ in range(max_workers)]

    def run(B):
        W = min(B, max_workers)
        def work(w):
            for i in range(w, B, W):
                adapt.process_audio_adaptive_ptr(ctxs[w], inputs[i % len(inputs)].data_ptr(), n, RATE, 1, gpudsp.FMT_FLT,
                                                 outs[w].data_ptr(), out_cap, True)
        ts = [threading.Thread(target=work, args=(w,)) for w in range(W)]
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        torch.cuda.synchronize()
        return time.perf_counter() - t0

    run(max_workers)                                   # warm every context (arena growth, table uploads)
    res = {}
    B = 1
    while B <= steps_cap:
        dt = run(B)
        res[str(B)] = {"seconds": dt, "realtime_x": B * 600.0 / dt, "workers": min(B, max_workers)}
        B *= 2
    for c in ctxs:
        c.close()
    del inputs, outs
    return {"streams": "10 min 48 kHz mono f32 conversational, ProcessAudio (adaptive), device resident", "by_batch": res,
            "note": "one worker thread + jt_ctx per concurrent stream, at most 8; wall clock between device synchronisations"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--minutes", type=int, default=MINUTES)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip fixed_spec / flac_encode / batch_sweep / stream_sharded")
    ap.add_argument("--sharded-hours", type=float, default=3.0)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    sys.path.insert(0, ROOT)
    import numpy as np
    import torch
    import torch.distributed as dist
    from jivetalking_b200 import adapt, gpudsp, shard

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the chain has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    my_files = shard.assign_files(world, rank, world)               # one file per GPU (configs[2])
    x = make_input(shard.stream_seed(12345, my_files[0]), args.minutes)   # C2 / C3: seeds 12345 + file index
    n = len(x)
    h_in = torch.from_numpy(x).pin_memory()
    d_in = h_in.cuda(non_blocking=False)
    out_cap = int(n * 44100 / RATE) + 3 * 4096
    d_out = torch.empty(out_cap, dtype=torch.int16, device="cuda")
    h_out = torch.empty(out_cap, dtype=torch.int16).pin_memory()
    ctx = gpudsp.Context(local)
    if world > 1 and not os.environ.get("JT_HOST_THREADS"):
        # the ranks of one node share its cores, and the barrier starts their steps together: each call's per-frame metadata
        # workers get a share of the cores instead of eight threads per rank piling onto all of them
        try:
            cores = len(os.sched_getaffinity(0))
        except AttributeError:
            cores = os.cpu_count() or 1
        gpudsp.lib().jt_set_host_threads(max(2, min(8, cores // world)))
    lib_stream = torch.cuda.ExternalStream(ctx.cuda_stream(), device=torch.device("cuda", local))   # the stream the kernels run on

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_dev():
        return adapt.process_audio_adaptive_ptr(ctx, d_in.data_ptr(), n, RATE, 1, gpudsp.FMT_FLT, d_out.data_ptr(), out_cap, True)

    # e2e: a worker that goes through file after file double-buffers its input (jt_prefetch_input): while step k computes, the
    # pinned host PCM of step k + 1 is on its way.  Every step still uploads its 691 MB and downloads its 318 MB inside the timed
    # region; two pinned host buffers alternate so that consecutive steps hand the library different "files".
    h_ins = [h_in, torch.from_numpy(x).pin_memory()]

    def run_e2e(k_steps):
        ctx.prefetch_input_ptr(h_ins[0].data_ptr(), n, 1, gpudsp.FMT_FLT)
        r = None
        for k in range(k_steps):
            if k + 1 < k_steps:
                ctx.prefetch_input_ptr(h_ins[(k + 1) & 1].data_ptr(), n, 1, gpudsp.FMT_FLT)
            r = adapt.process_audio_adaptive_ptr(ctx, h_ins[k & 1].data_ptr(), n, RATE, 1, gpudsp.FMT_FLT, h_out.data_ptr(), out_cap, False)
        return r

    def step_e2e():
        return adapt.process_audio_adaptive_ptr(ctx, h_in.data_ptr(), n, RATE, 1, gpudsp.FMT_FLT, h_out.data_ptr(), out_cap, False)

    # ---- device-resident timing (value): no per-kernel instrumentation ---------------------------
    for _ in range(args.warmup):
        res, an = step_dev()
    ctx.reset_counters()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        t0 = time.perf_counter()
        ev0.record(lib_stream)
        for _ in range(args.steps):
            res, an = step_dev()
        ev1.record(lib_stream)
        torch.cuda.synchronize()
        barrier()
        t_wall = time.perf_counter() - t0
    # device clock (CUDA events on the library's own stream) around the K steps, host gaps between kernels included
    t_dev = ev0.elapsed_time(ev1) * 1e-3
    launches = ctx.launch_count()
    # ---- the same K steps once more with an event pair around every kernel group: the per-kernel breakdown and the device-idle
    #      gaps (2 events per group cost ~1 % of a step, so this pass is not the one `value` is taken from) -------------------
    ctx.enable_timing(True)
    step_dev()
    ctx.reset_counters()
    ev0.record(lib_stream)
    for _ in range(args.steps):
        step_dev()
    ev1.record(lib_stream)
    torch.cuda.synchronize()
    t_dev_instr = ev0.elapsed_time(ev1) * 1e-3
    timings = ctx.kernel_timings()
    ctx.enable_timing(False)

    # ---- end to end through the C ABI with host buffers (e2e) ----------------------------------
    run_e2e(max(2, args.warmup // 2))
    barrier()
    t0 = time.perf_counter()
    ev0.record(lib_stream)
    res_e, _ = run_e2e(args.steps)
    ev1.record(lib_stream)
    barrier()
    t_e2e_wall = time.perf_counter() - t0
    # the same K steps as single blocking calls (upload, compute and download of one file strictly one after the other)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    t_e2e_single = time.perf_counter() - t0
    # the call returns after its device->host copy has landed, so the wall clock between the barriers IS the end-to-end time;
    # the event pair on the library stream is reported next to it
    t_e2e = max(t_e2e_wall, ev0.elapsed_time(ev1) * 1e-3)

    extras = {}
    if rank == 0 and not args.no_extras:
        # ---- the caller-supplied-spec entry (DefaultFilterConfig spec): Pass 2 is enqueued before Pass 1 -------------------
        for _ in range(2):
            ctx.process_audio_ptr(d_in.data_ptr(), n, RATE, 1, gpudsp.FMT_FLT, d_out.data_ptr(), out_cap, True)
        torch.cuda.synchronize()
        ev0.record(lib_stream)
        for _ in range(args.steps):
            res_f = ctx.process_audio_ptr(d_in.data_ptr(), n, RATE, 1, gpudsp.FMT_FLT, d_out.data_ptr(), out_cap, True)
        ev1.record(lib_stream)
        torch.cuda.synchronize()
        t_fx = ev0.elapsed_time(ev1) * 1e-3 / args.steps
        extras["fixed_spec"] = {"value": n / t_fx, "unit": "samples/s", "realtime_x": n / t_fx / RATE, "ms_per_step": 1e3 * t_fx,
                                "entry": "jt_process_audio_dev with DefaultFilterConfig's spec (filters.go:353-355), same input"}
        # ---- the result's container: FLAC stream of the s16 44.1 kHz output, device resident -------------------
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
        extras["flac_encode"] = {"ms_per_step": 1e3 * t_fl, "realtime_x": (n_out / 44100) / t_fl, "samples_per_s": n_out / t_fl,
                                 "bytes": int(nbytes), "ratio": nbytes / (2.0 * n_out), "algorithmic_GBps": (2.0 * n_out + nbytes) / t_fl / 1e9,
                                 "entry": "jt_flac_encode_dev (4096-sample frames; encoder.go:92-101)"}
        del d_flac
    barrier()
    if world == 1 and not args.no_extras:
        del d_in, d_out
        torch.cuda.empty_cache()
        extras["batch_sweep"] = batch_sweep(local, torch, gpudsp, adapt)
        d_in = d_out = None
    if world > 1 and not args.no_extras and hasattr(shard, "bench_stream_sharded"):
        del d_in, d_out
        torch.cuda.empty_cache()
        sh = shard.bench_stream_sharded(ctx, local, rank, world, hours=args.sharded_hours)
        if rank == 0 and sh is not None:
            extras["stream_sharded"] = sh
        d_in = d_out = None
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
    traffic, traffic_src = ncu_traffic()
    timings.sort(key=lambda t: -t[1])
    gaps = [t for t in timings if t[0].startswith("gap:")]
    timings = [t for t in timings if not t[0].startswith("gap:")]
    gpu_timings = [t for t in timings if not t[0].startswith("host:")]
    dom = gpu_timings[0] if gpu_timings else ("none", 0.0, 0)
    dom_ms_per_step = dom[1] / args.steps
    dom_bytes = KERNEL_BYTES.get(dom[0].split(":")[0], 8.0) * n
    achieved = dom_bytes / (dom_ms_per_step * 1e-3) / 1e9 if dom_ms_per_step > 0 else 0.0
    kernel_ms = sum(t[1] for t in gpu_timings) / args.steps
    va = an.voice_activity
    line = {
        "metric": "samples/s full 4-pass chain", "value": value, "unit": "samples/s", "realtime_x": value / RATE,
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_dev / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32/f64", "data": "synthetic",
        "config": {"workload": f"ProcessAudio (adaptive), {args.minutes} min 48 kHz mono f32 conversational stream per GPU "
                               "(BASELINE.json configs[1]; file per GPU = configs[2])",
                   "entry": "jt_process_audio_adaptive[_dev]: Pass 1, detector, 17 band graphs, AdaptConfig, Pass 2, region re-measures, Pass 3, Pass 4, region re-measures",
                   "samples_per_gpu": n, "pass2_spec": an.pass2_spec.decode(),
                   "l2": "inputs (691 MB/stream) and every intermediate exceed the 126 MB L2"},
        "e2e": {"value": e2e, "unit": "samples/s", "realtime_x": e2e / RATE, "h2d_bytes_per_step": int(n * 4),
                "d2h_bytes_per_step": int(res_e.n_out * 2), "ms_per_step": 1e3 * t_e2e / args.steps,
                "mode": "jt_prefetch_input(file k+1) + jt_process_audio_adaptive(file k), pinned host buffers: each step's upload overlaps the previous step's kernels; the first upload of the timed region is exposed",
                "single_call_ms_per_step": 1e3 * t_e2e_single / args.steps, "single_call_value": total_samples / t_e2e_single},
        "gpu_launches": int(launches),
        "timing": {"method": "CUDA events on the library stream around the K steps (max over ranks); e2e: wall clock between barriers, the call returns after its D2H copy",
                   "wall_ms_per_step": 1e3 * t_wall / args.steps, "e2e_wall_ms_per_step": 1e3 * t_e2e_wall / args.steps,
                   "instrumented_ms_per_step": 1e3 * t_dev_instr / args.steps,
                   "breakdown": "kernels_ms_per_step / device_idle_ms_per_step come from a second pass of K steps with an event pair around every kernel group"},
        "roofline": {"bound": "hbm", "kernel": dom[0], "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak if peak else None,
                     "traffic": traffic.get(dom[0]) if args.minutes == 60 else None,
                     "traffic_source": traffic_src, "peak_source": peak_src,
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
                   "n_out": int(res.n_out), "limiter_needed": int(res.limiter_needed), "pass4_type": int(res.pass4.normalization_type),
                   "speech_profile": bool(va.has_speech_profile), "noise_profile": bool(va.has_noise_profile),
                   "voice_activated": bool(va.voice_activated), "noise_floor": va.floor,
                   "regions_remeasured": int(an.filtered_regions.has_room_tone + an.filtered_regions.has_speech +
                                             an.final_regions.has_room_tone + an.final_regions.has_speech)},
    }
    line.update(extras)
    if not args.no_cpu_baseline and world == 1:
        cores = os.cpu_count() or 1
        v, total, busy = cpu_baseline(cores, CPU_SAMPLE_SECONDS)
        line["cpu_baseline"] = {"value": v, "unit": "samples/s", "realtime_x": v / RATE, "cores": cores, "kind": "port",
                                "sample": f"{cores} x {CPU_SAMPLE_SECONDS:.0f} s conversational streams, one scalar oracle ProcessAudio per core "
                                          f"({busy:.1f} s of CPU work per core)"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
