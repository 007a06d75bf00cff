#!/usr/bin/env python
"""BASELINE.json configs[3]: ONE long stream (default 3 h, 96 kHz, stereo, f32) through the four passes with every
pass cut into one chunk per GPU (jt_analyse_chunk / jt_graph_chunk + host merges, shard.process_stream_sharded).

  python scripts/stream_shard_bench.py --hours 3                               # one GPU: the whole stream on it
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
         scripts/stream_shard_bench.py --hours 3                               # one chunk per GPU

Timing: barrier + device synchronize on both sides of each repetition, max over ranks.  The timed region starts with
the stream in HOST memory on every rank (a production caller decodes only its window) and ends with the 44.1 kHz s16
result gathered on every rank; it contains the host->device copies of every window, the exchange of afftdn's noise
floor carry, the all-gathers of the measurement blobs and of the Pass-2 / Pass-4 audio, and the host merges.
Prints one JSON line on rank 0.  The stream is a 10 min segment of the C2 recipe tiled to the requested length
(R = L delayed 7 samples x 0.9), so no rank spends minutes synthesising."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
from jivetalking_b200 import gpudsp, shard, synth


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--hours", type=float, default=3.0)
    ap.add_argument("--rate", type=int, default=96000)
    ap.add_argument("--channels", type=int, default=2)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--check", action="store_true", help="rank 0 also runs the whole stream on one GPU and compares")
    ap.add_argument("--adaptive", action="store_true",
                    help="conversational input; every rank derives the Pass-2 spec from the sharded Pass 1 (detector + AdaptConfig)")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    seg = synth.podcast_like(600.0, a.rate, seed=12345) if a.adaptive else synth.speech_like(600.0, a.rate, seed=12345)
    n = int(a.hours * 3600 * a.rate)
    # the stream lives in PINNED host memory (what a decoder thread would fill): windows go to the GPU at PCIe rate
    pcm = torch.empty(n * a.channels, dtype=torch.float32, pin_memory=True).numpy()
    st = synth.stereo_from_mono(seg) if a.channels == 2 else seg
    for o in range(0, pcm.size, st.size):
        m = min(st.size, pcm.size - o)
        pcm[o: o + m] = st[:m]
    del st
    comm = shard.DistComm(dev)
    times = []
    with gpudsp.Context(local) as ctx:
        for rep in range(a.reps + 1):                      # first repetition = warm-up (arena growth, table uploads)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            phases = {}
            if a.adaptive:
                out, r = shard.process_stream_sharded_adaptive(ctx, comm, pcm, a.rate, a.channels, timings=phases, device=dev)
                m1, iv = r["input"], r["intervals"]
            else:
                m1, iv = shard.analyse_stream_sharded(ctx, pcm, a.rate, a.channels, device=dev)
                phases["pass1"] = time.perf_counter() - t0
                out, r = shard.process_stream_sharded(ctx, comm, pcm, a.rate, a.channels, timings=phases)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            dt = time.perf_counter() - t0
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            if rep > 0:
                times.append(float(t[0]))
        check = None
        if a.check and rank == 0:
            ctx.process_audio(pcm[: 10 * a.rate * a.channels], a.rate, a.channels)
            t0 = time.perf_counter()
            if a.adaptive:
                from jivetalking_b200 import adapt
                pcm1, res1, an1 = adapt.process_audio_adaptive(ctx, pcm, a.rate, a.channels)
            else:
                pcm1, res1 = ctx.process_audio(pcm, a.rate, a.channels)
            single = time.perf_counter() - t0
            d = (out.astype(np.int32) - pcm1.astype(np.int32)) / 32768.0
            check = {"pcm_rms_diff": float(np.sqrt(np.mean(d * d))), "n_out_equal": bool(len(out) == len(pcm1)),
                     "d_final_lufs": r["final"].input_i - res1.final.input_i, "d_final_dbtp": r["final"].input_tp - res1.final.input_tp,
                     "d_final_lra": r["final"].input_lra - res1.final.input_lra, "d_input_lufs": m1.input_i - res1.input.input_i, "single_gpu_jt_process_audio_seconds": single}
            if a.adaptive:
                check["spec_equal"] = bool(r["specs"][0] == an1.pass2_spec.decode())
    if rank == 0:
        best = min(times)
        print(json.dumps({
            "metric": "x realtime, one stream sharded over GPUs, full 4-pass chain (host buffers, gathers and merges inside)",
            "config": {"workload": f"single {a.hours:g} h {a.rate} Hz {a.channels}-channel f32 stream, one chunk per GPU per pass (BASELINE.json configs[3])"},
            "n_gpus": world, "seconds_per_stream": best, "rank0_phase_seconds_last_rep": {k: round(v, 4) for k, v in phases.items()}, "all_seconds": times,
            "realtime_x": a.hours * 3600 / best, "frames_per_s": n / best,
            "input_lufs": m1.input_i, "input_dbtp": m1.input_tp, "final_lufs": r["final"].input_i, "final_dbtp": r["final"].input_tp,
            "final_lra": r["final"].input_lra, "n_out": int(len(out)), "pass3_input_i": r["pass3"].input_i, "check": check,
            "adaptive": bool(a.adaptive), "pass2_spec": r["specs"][0] if a.adaptive else "DefaultFilterConfig"}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
