#!/usr/bin/env python
"""BASELINE.json configs[4]: B concurrent 48 kHz mono streams through the full four-pass chain on ONE GPU.
Each stream has its own jt_ctx (= its own CUDA stream, like one worker goroutine per file,
cmd/jivetalking/pool.go:122-153) driven by its own host thread; streams are device-resident.
Prints one JSON line per batch size: aggregate samples/s, x realtime, chain HBM-roofline fraction.

Usage: python scripts/batch_sweep.py [--minutes 10] [--batches 1,2,4,8,16,32] [--steps 2]
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (make_input, measured_peaks, constants)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--minutes", type=int, default=10)
    ap.add_argument("--batches", default="1,2,4,8,16,32")
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--device", type=int, default=0)
    args = ap.parse_args()
    import torch
    from jivetalking_b200 import gpudsp
    torch.cuda.set_device(args.device)
    batches = [int(b) for b in args.batches.split(",")]
    bmax = max(batches)
    distinct = min(bmax, 8)                       # distinct seeds; larger batches reuse the inputs (read-only)
    xs = [torch.from_numpy(bench.make_input(12345 + i, args.minutes)).cuda() for i in range(distinct)]
    n = xs[0].numel()
    cap = int(n * 44100 / bench.RATE) + 3 * 4096
    peak, _ = bench.measured_peaks()
    ctxs, outs = [], []
    for b in batches:
        while len(ctxs) < b:
            ctxs.append(gpudsp.Context(args.device))
            outs.append(torch.empty(cap, dtype=torch.int16, device="cuda"))

        def work(i, reps):
            for _ in range(reps):
                ctxs[i].process_audio_ptr(xs[i % distinct].data_ptr(), n, bench.RATE, 1, gpudsp.FMT_FLT,
                                          outs[i].data_ptr(), cap, True)

        def run(reps):
            th = [threading.Thread(target=work, args=(i, reps)) for i in range(b)]
            t0 = time.perf_counter()
            for t in th:
                t.start()
            for t in th:
                t.join()
            torch.cuda.synchronize()
            return time.perf_counter() - t0

        run(1)                                    # warm-up (pool growth, plan caches)
        dt = run(args.steps)
        sps = b * n * args.steps / dt
        print(json.dumps({"batch": b, "minutes_per_stream": args.minutes, "steps": args.steps, "wall_s": round(dt, 4),
                          "samples_per_s": sps, "realtime_x": sps / bench.RATE,
                          "chain_hbm_frac": bench.BYTES_PER_SAMPLE_4PASS * sps / 1e9 / peak,
                          "mem_gb": round(torch.cuda.mem_get_info()[1] / 1e9 - torch.cuda.mem_get_info()[0] / 1e9, 1)}),
              flush=True)


if __name__ == "__main__":
    main()
