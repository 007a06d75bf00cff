import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from jivetalking_b200 import adapt, gpudsp
minutes = int(sys.argv[1]) if len(sys.argv) > 1 else 60


def main():
    x = bench.make_input(12345, minutes); n = len(x)
    h_in = torch.from_numpy(x).pin_memory(); h_in2 = torch.from_numpy(x).pin_memory(); d_in = h_in.cuda()
    cap = int(n * 44100 / 48000) + 3 * 4096
    d_out = torch.empty(cap, dtype=torch.int16, device="cuda"); h_out = torch.empty(cap, dtype=torch.int16).pin_memory()
    ctx = gpudsp.Context(0)
    def sync(): torch.cuda.synchronize()
    def t(fn, k=3):
        fn(); sync(); t0 = time.perf_counter()
        for _ in range(k): fn()
        sync(); return (time.perf_counter() - t0) / k * 1e3
    dev = lambda: adapt.process_audio_adaptive_ptr(ctx, d_in.data_ptr(), n, 48000, 1, gpudsp.FMT_FLT, d_out.data_ptr(), cap, True)
    host = lambda: adapt.process_audio_adaptive_ptr(ctx, h_in.data_ptr(), n, 48000, 1, gpudsp.FMT_FLT, h_out.data_ptr(), cap, False)
    def pre():
        ctx.prefetch_input_ptr(h_in.data_ptr(), n, 1, gpudsp.FMT_FLT); sync()
        t0 = time.perf_counter()
        adapt.process_audio_adaptive_ptr(ctx, h_in.data_ptr(), n, 48000, 1, gpudsp.FMT_FLT, h_out.data_ptr(), cap, False)
        return (time.perf_counter() - t0) * 1e3
    print("dev/dev       %.2f ms" % t(dev))
    print("host/host     %.2f ms" % t(host))
    pre(); print("resident in / host out  %.2f ms" % np.mean([pre() for _ in range(3)]))
    print("H2D 691 MB    %.2f ms" % t(lambda: d_in.copy_(h_in, non_blocking=True)))
    print("D2H 318 MB    %.2f ms" % t(lambda: h_out.copy_(d_out, non_blocking=True)))
    hs = [h_in, h_in2]
    ctx.prefetch_input_ptr(hs[0].data_ptr(), n, 1, gpudsp.FMT_FLT)
    marks = [time.perf_counter()]
    for k in range(6):
        if k + 1 < 6:
            ctx.prefetch_input_ptr(hs[(k + 1) & 1].data_ptr(), n, 1, gpudsp.FMT_FLT)
        marks.append(time.perf_counter())
        adapt.process_audio_adaptive_ptr(ctx, hs[k & 1].data_ptr(), n, 48000, 1, gpudsp.FMT_FLT, h_out.data_ptr(), cap, False)
        marks.append(time.perf_counter())
    print("pipelined: prefetch-call / process-call ms:", [round((marks[i + 1] - marks[i]) * 1e3, 2) for i in range(len(marks) - 1)])
    def timed(with_prefetch):
        ctx.prefetch_input_ptr(hs[0].data_ptr(), n, 1, gpudsp.FMT_FLT); sync()
        ctx.enable_timing(True); ctx.reset_counters()
        if with_prefetch:
            ctx.prefetch_input_ptr(hs[1].data_ptr(), n, 1, gpudsp.FMT_FLT)
        t0 = time.perf_counter()
        adapt.process_audio_adaptive_ptr(ctx, hs[0].data_ptr(), n, 48000, 1, gpudsp.FMT_FLT, h_out.data_ptr(), cap, False)
        dt = (time.perf_counter() - t0) * 1e3
        sync()
        tk = {k: ms for k, ms, nl in ctx.kernel_timings()}
        ctx.enable_timing(False)
        return dt, tk
    timed(False)
    d0, k0 = timed(False); d1, k1 = timed(True)
    print("timed: resident %.2f ms, with a concurrent prefetch %.2f ms" % (d0, d1))
    print("differences > 0.25 ms:", {k: (round(k0.get(k, 0), 2), round(k1.get(k, 0), 2)) for k in sorted(set(k0) | set(k1)) if abs(k0.get(k, 0) - k1.get(k, 0)) > 0.25})
    ctx.enable_timing(True); host(); ctx.reset_counters(); host(); sync()
    tk = {k: ms for k, ms, nl in ctx.kernel_timings()}
    print("gaps (host/host):", {k: round(v, 2) for k, v in tk.items() if k.startswith("gap:") and v > 0.2})


if __name__ == "__main__":
    main()
