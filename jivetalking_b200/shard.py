"""Multi-GPU structure of the path: files per GPU (BASELINE.json configs[2]) and, for one long stream,
contiguous chunks per GPU with ONE all-gather of mergeable analysis values (configs[3]).

The reference's data parallelism is one goroutine per file behind a semaphore
(cmd/jivetalking/pool.go:122-153); here it is one process per GPU, files dealt round-robin, no
data-path collective.  torch.distributed (NCCL on GPUs, gloo in the CPU tests) carries only the
barrier and the max-over-ranks timing the bench contract asks for."""
import torch
import torch.distributed as dist


def assign_files(n_files, rank, world):
    """Indices of the files rank `rank` processes (round-robin, like a worker pool draining a queue)."""
    return list(range(rank, n_files, world))


def stream_seed(base_seed, file_index):
    """C3: 8 copies of the C2 recipe with seeds 12345..12352."""
    return base_seed + file_index


def job_throughput(samples_local, seconds_local, device=None):
    """Whole-job samples/s: sum of samples over ranks / max of times over ranks."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return samples_local / seconds_local, samples_local, seconds_local
    t = torch.tensor([float(seconds_local)], dtype=torch.float64, device=device)
    s = torch.tensor([float(samples_local)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(s, op=dist.ReduceOp.SUM)
    return float(s[0]) / float(t[0]), float(s[0]), float(t[0])


# ---- one long stream over several GPUs (configs[3]) ------------------------------------------------
def plan_stream_chunks(total_frames, unit, world):
    """Contiguous chunks [(first, owned)] tiling [0, total_frames): every boundary a multiple of `unit`
    (jt_analyse_chunk_unit: whole ticks / decoder frames / spectral hops), sizes balanced to within one
    unit, the last chunk takes the partial tail.  Ranks beyond the number of units get (total, 0)."""
    units = max(1, -(-total_frames // unit))
    per, extra = divmod(units, world)
    chunks, first = [], 0
    for r in range(world):
        n_units = per + (1 if r < extra else 0)
        owned = min(n_units * unit, max(total_frames - first, 0))
        chunks.append((first, owned))
        first += owned
    return chunks


def local_range(first, owned, total_frames, unit):
    """The frames a rank must hold to analyse its chunk: one unit of left context in mid-stream, one unit
    (or up to the stream's end) of right context."""
    lo = max(0, first - unit)
    hi = min(total_frames, first + owned + unit)
    return lo, hi


def allgather_blobs(blob, device=None):
    """Every rank's byte blob on every rank: one all-gather of padded uint8 tensors (NCCL on GPUs, gloo on
    CPU) plus one of the lengths.  Returns a list of bytes objects in rank order."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [bytes(blob)]
    world = dist.get_world_size()
    n = torch.tensor([len(blob)], dtype=torch.int64, device=device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    cap = int(max(int(x[0]) for x in sizes))
    buf = torch.zeros(cap, dtype=torch.uint8, device=device)
    if len(blob):
        buf[: len(blob)] = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(buf.device)
    parts = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf)
    return [bytes(parts[r][: int(sizes[r][0])].cpu().numpy().tobytes()) for r in range(world)]


def analyse_stream_sharded(ctx, pcm, rate, channels=1, device=None):
    """Pass-1 analysis of ONE stream by all ranks of the process group.  `pcm` is the whole stream here (tests,
    bench); a production caller decodes only local_range().  Returns (Measurements, [Interval]) on every rank."""
    from . import gpudsp
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    total = pcm.size // channels
    unit = gpudsp.analyse_chunk_unit(rate)
    first, owned = plan_stream_chunks(total, unit, world)[rank]
    blob = b""
    if owned > 0:
        lo, hi = local_range(first, owned, total, unit)
        blob = ctx.analyse_chunk(pcm.reshape(-1)[lo * channels: hi * channels], rate, channels, lo, first, owned, total)
    blobs = [b for b in allgather_blobs(blob, device=device) if len(b)]
    return gpudsp.analyse_merge(blobs, total, rate)


# ---- Passes 2-4 of one stream over several GPUs (configs[3]) ------------------------------------------
# Every pass is a graph (a spec string) run on contiguous chunks with context (jt_graph_chunk); what crosses rank
# boundaries is (1) afftdn's noise-floor carry, inside the pass, through the exchange callback, (2) one all-gather of
# the chunks' measurement blobs per measuring pass, merged on every rank (jt_graph_merge), and (3) the Pass-2 output
# itself (the reference's intermediate FLAC), gathered so that Pass 3/4 can cut their own chunk grid at 44.1 kHz.
class LocalComm:
    """All chunks in ONE process, run in stream order (tests on one GPU, and the reference for what the ranks of a
    process group do collectively)."""

    def __init__(self, n_chunks):
        self.world = n_chunks
        self._carries = []

    def ranks(self):
        return range(self.world)

    def begin_pass(self):
        self._carries = []

    def exchange(self, send):
        # chunk k enters after chunks 0..k-1: their records are known, later chunks' are not needed (the library
        # only composes records with a smaller stream position)
        self._carries.append(bytes(send))
        return b"".join(self._carries) + bytes(len(send)) * (self.world - len(self._carries))

    def gather_blobs(self, per_rank):
        return [b for b in per_rank if len(b)]

    pinned_results = False

    def gather_pcm(self, per_rank):
        import numpy as np
        return np.concatenate([p for p in per_rank if p is not None and len(p)])


class DistComm:
    """One chunk per rank of the torch.distributed process group (NCCL on GPUs, gloo in the CPU tests)."""

    def __init__(self, device=None):
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.device = device

    def ranks(self):
        return [self.rank]

    def begin_pass(self):
        pass

    def exchange(self, send):
        if self.world == 1:
            return bytes(send)
        t = torch.frombuffer(bytearray(send), dtype=torch.uint8)
        if self.device is not None:
            t = t.to(self.device)
        parts = [torch.zeros_like(t) for _ in range(self.world)]
        dist.all_gather(parts, t)
        return b"".join(p.cpu().numpy().tobytes() for p in parts)

    def exchange_raw(self, send_addr, nbytes, recv_addr):
        """the same all-gather on the library's own buffers (jt_set_exchange with raw=True): no Python-level copies"""
        import ctypes as C
        src = torch.frombuffer((C.c_ubyte * nbytes).from_address(send_addr), dtype=torch.uint8)
        dst = torch.frombuffer((C.c_ubyte * (nbytes * self.world)).from_address(recv_addr), dtype=torch.uint8)
        if self.world == 1:
            dst.copy_(src)
            return
        if self.device is not None:
            out = torch.empty(nbytes * self.world, dtype=torch.uint8, device=self.device)
            dist.all_gather_into_tensor(out, src.to(self.device))
            dst.copy_(out)
        else:
            parts = [torch.empty(nbytes, dtype=torch.uint8) for _ in range(self.world)]
            dist.all_gather(parts, src)
            for r, p in enumerate(parts):
                dst[r * nbytes:(r + 1) * nbytes].copy_(p)

    def gather_blobs(self, per_rank):
        return [b for b in allgather_blobs(per_rank[0], device=self.device) if len(b)]

    pinned_results = True          # one chunk per pass and rank: it can stay in the context's pinned buffer until gathered

    def gather_pcm(self, per_rank):
        """Ragged all-gather of every rank's owned samples, in rank (= stream) order: bytes on the wire (gloo has no
        int16 collectives, NCCL does not care), one padded all-gather on the device, one device->host copy per part
        straight into a pinned result."""
        import numpy as np
        a = np.ascontiguousarray(per_rank[0])
        if self.world == 1:
            return a.copy() if self.pinned_results else a
        raw = a.view(np.uint8).reshape(-1)
        n = torch.tensor([raw.size], dtype=torch.int64, device=self.device)
        sizes = [torch.zeros_like(n) for _ in range(self.world)]
        dist.all_gather(sizes, n)
        sizes = [int(x[0]) for x in sizes]
        cap = max(max(sizes), 1)
        buf = torch.empty(cap, dtype=torch.uint8, device=self.device)
        if raw.size:
            buf[: raw.size].copy_(torch.from_numpy(raw), non_blocking=True)
        allp = torch.empty(cap * self.world, dtype=torch.uint8, device=self.device)
        dist.all_gather_into_tensor(allp, buf)
        on_gpu = allp.is_cuda
        res = torch.empty(sum(sizes), dtype=torch.uint8, pin_memory=on_gpu)
        off = 0
        for r in range(self.world):
            res[off: off + sizes[r]].copy_(allp[r * cap: r * cap + sizes[r]], non_blocking=on_gpu)
            off += sizes[r]
        if on_gpu:
            torch.cuda.synchronize(allp.device)
        self._last_result = res            # keeps the pinned allocation alive as long as the caller may hold the view
        return res.numpy().view(a.dtype)


def run_graph_sharded(ctx, comm, spec, pcm, rate, channels=1, want_pcm=True, context=None, timings=None, tag="",
                      want_meta=False):
    """One graph over one stream, one chunk per rank of `comm`.  Returns (sink pcm of the whole stream or None,
    merged dict(meta, loudnorm, measurements)) on every rank; meta (the sink-frame records) only with want_meta."""
    from . import gpudsp
    total = pcm.size // channels
    unit = gpudsp.graph_chunk_unit(spec, rate)
    left, right = context if context else gpudsp.graph_chunk_context(spec, rate)
    chunks = plan_stream_chunks(total, unit, comm.world)
    n_exch = gpudsp.graph_exchanges(spec)
    import time
    t0 = time.perf_counter()
    ctx.set_exchange(comm.exchange, comm.world)
    comm.begin_pass()
    blobs, parts = [], []
    try:
        for r in comm.ranks():
            first, owned = chunks[r]
            if owned <= 0:                     # more ranks than units: stay in step with the collectives
                for _ in range(n_exch):
                    comm.exchange(bytes(32))
                blobs.append(b"")
                parts.append(None)
                continue
            lo = max(0, first - left)
            hi = min(total, first + owned + right)
            res = ctx.graph_chunk(spec, pcm.reshape(-1)[lo * channels: hi * channels], rate, channels, lo, first, owned,
                                  total, want_pcm=want_pcm, pinned=comm.pinned_results)
            blobs.append(res["blob"])
            parts.append(res["pcm"])
    finally:
        ctx.set_exchange(None, 1)
    t1 = time.perf_counter()
    fmt = gpudsp._FMT_OF_NP[pcm.dtype]
    merged = gpudsp.graph_merge(spec, comm.gather_blobs(blobs), total, rate, channels, fmt, want_meta=want_meta)
    t2 = time.perf_counter()
    out = None
    if want_pcm:
        import numpy as np
        dt = next(p.dtype for p in parts if p is not None)
        out = comm.gather_pcm([p if p is not None else np.zeros(0, dtype=dt) for p in parts])
    if timings is not None:
        timings[tag + ":chunks"] = t1 - t0
        timings[tag + ":blob_gather_merge"] = t2 - t1
        timings[tag + ":pcm_gather"] = time.perf_counter() - t2
    return out, merged


def process_stream_sharded(ctx, comm, pcm, rate, channels=1, pass2_spec=None,
                           target_i=-16.0, target_tp=-1.0, target_lra=20.0, timings=None):
    """The four-pass chain (ProcessAudio, processor.go:78-216) of ONE stream over the ranks of `comm`: what
    jt_process_audio does on one GPU, with every pass cut into chunks.  Returns (int16 mono 44.1 kHz output,
    dict(filtered, final, pass3, pass4, plan, specs)) on every rank.  Pass 1 is analyse_stream_sharded."""
    from . import gpudsp
    spec2 = pass2_spec or gpudsp.default_pass2_spec()
    out2, mg2 = run_graph_sharded(ctx, comm, spec2, pcm, rate, channels, want_pcm=True, timings=timings, tag="pass2")
    filtered = mg2["measurements"]
    spec3, plan = gpudsp.build_pass3_spec(filtered.input_i, filtered.input_tp, target_i, target_tp, target_lra)
    _, mg3 = run_graph_sharded(ctx, comm, spec3, out2, 44100, 1, want_pcm=False, timings=timings, tag="pass3")
    p3 = mg3["loudnorm"]
    if not (p3.input_i > -70.0):
        raise gpudsp.JtError(-1, f"cannot normalise silent audio (measured {p3.input_i:.1f} LUFS)")
    spec4, eff, off = gpudsp.build_pass4_spec(plan, p3, target_i, target_tp, target_lra, 44100)
    out4, mg4 = run_graph_sharded(ctx, comm, spec4, out2, 44100, 1, want_pcm=True, timings=timings, tag="pass4")
    return out4, dict(filtered=filtered, final=mg4["measurements"], pass3=p3, pass4=mg4["loudnorm"], plan=plan,
                      effective_target_i=eff, offset_db=off, specs=(spec2, spec3, spec4), pass2_pcm=out2)


def adapt_stream_sharded(ctx, pcm, rate, channels=1, base=None, device=None):
    """AnalyseAudio + AdaptConfig (analyser.go:325-372, adaptive.go:13-40) for ONE stream analysed by all ranks: Pass 1 is
    analyse_stream_sharded (one all-gather); the merged measurements and intervals are identical on every rank, so every
    rank runs the same deterministic detector and AdaptConfig and arrives at the same Pass-2 spec without further
    communication.  The 17 band graphs read only the elected regions (<= 60 s of speech, <= 18 s of room tone
    -- the reference re-decodes them 17 times, analyser_bands.go:106-167): each rank measures them itself from the
    region's samples.  Returns (Measurements, [Interval], VoiceActivity, FilterConfig, AdaptDiagnostics, spec)."""
    from . import adapt
    m, iv = analyse_stream_sharded(ctx, pcm, rate, channels, device=device)
    va, _, _ = adapt.detect_voice_activity(m, iv)
    lo, hi = adapt.band_plan()
    flat = pcm.reshape(-1)

    def region_bands(start_ns, dur_ns, lo_hz, hi_hz):
        # atrim=start=%f:duration=%f (analyser_bands.go:54-60): the same sample window jt_band_rms cuts from the whole stream
        st, du = float("%f" % (start_ns / 1e9)), float("%f" % (dur_ns / 1e9))
        s0 = (round(st * 1e6) * rate + 500000) // 1000000
        n = (round(du * 1e6) * rate + 500000) // 1000000
        a, b = min(max(s0, 0), flat.size // channels), min(flat.size // channels, s0 + n)
        return ctx.band_rms(flat[a * channels: b * channels], rate, 0.0, du, lo_hz, hi_hz, channels=channels)

    speech = noise = None
    if va.has_speech_profile and va.speech_profile.region.duration_ns > 0:
        r, f = region_bands(va.speech_profile.region.start_ns, va.speech_profile.region.duration_ns, lo[:2], hi[:2])
        speech = (list(r), list(f))
    if va.has_noise_profile and va.noise_profile.duration_ns > 0:
        r, f = region_bands(va.noise_profile.start_ns, va.noise_profile.duration_ns, lo[2:], hi[2:])
        noise = (list(r), list(f))
    adapt.apply_band_rms(va, speech, noise)
    cfg, diag = adapt.adapt_config(m, va, base)
    return m, iv, va, cfg, diag, adapt.build_filter_spec(cfg)


def process_stream_sharded_adaptive(ctx, comm, pcm, rate, channels=1, base=None, timings=None, device=None):
    """ProcessAudio (processor.go:78-216) of ONE stream over the ranks of `comm` with the adaptive Pass-2 spec derived on
    every rank from the sharded Pass 1 (adapt_stream_sharded)."""
    import time
    t0 = time.perf_counter()
    m, iv, va, cfg, diag, spec = adapt_stream_sharded(ctx, pcm, rate, channels, base=base, device=device)
    if timings is not None:
        timings["pass1+adapt"] = time.perf_counter() - t0
    out, info = process_stream_sharded(ctx, comm, pcm, rate, channels, pass2_spec=spec, target_i=cfg.loudnorm.target_i,
                                       target_tp=cfg.loudnorm.target_tp, target_lra=cfg.loudnorm.target_lra, timings=timings)
    info.update(input=m, intervals=iv, voice_activity=va, config=cfg, diagnostics=diag)
    return out, info


# ---- the whole ProcessAudio of one stream behind ONE call per rank (jt_process_audio_sharded) ----------------------------------
class ThreadComm:
    """All-gather between `world` threads of one process (tests on one GPU: every "rank" is a thread with its own jt_ctx)."""

    def __init__(self, world):
        import threading
        self.world = world
        self._slots = [b""] * world
        self._bar = threading.Barrier(world)

    def exchange_for(self, rank):
        def fn(send):
            self._slots[rank] = bytes(send)
            self._bar.wait()
            out = b"".join(self._slots)
            self._bar.wait()
            return out
        return fn


def process_stream_sharded_call(ctxs, pcm, rate, channels=1, adaptive=True, base=None):
    """ONE stream, len(ctxs) ranks as threads of this process, each calling jt_process_audio_sharded with its window.
    Returns (whole int16 output assembled from the owned parts, [per-rank (ProcessResult, Analysis, ShardTiming)])."""
    import threading
    import numpy as np
    from . import adapt
    world = len(ctxs)
    total = pcm.size // channels
    comm = ThreadComm(world)
    results, errors = [None] * world, [None] * world

    def work(r):
        try:
            p = adapt.sharded_plan(total, rate, world, r)
            win = pcm.reshape(-1)[p.local_first * channels: (p.local_first + p.n_local) * channels]
            results[r] = adapt.process_audio_sharded(ctxs[r], win, rate, channels, total, world, r, exchange=comm.exchange_for(r),
                                                     adaptive=adaptive, base=base)
        except Exception as e:          # noqa: BLE001  (a failing rank must not leave the others waiting at the barrier)
            errors[r] = e
            comm._bar.abort()

    ts = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    for e in errors:
        if e is not None and not isinstance(e, threading.BrokenBarrierError):
            raise e
    for e in errors:
        if e is not None:
            raise e
    n_out = int(results[0][2].n_out)
    out = np.zeros(n_out, dtype=np.int16)
    for own, first, res, an, tm in results:
        out[first: first + len(own)] = own
    return out, [(r[2], r[3], r[4]) for r in results]


def bench_stream_sharded(ctx, local, rank, world, hours=3.0, rate=96000, channels=2, reps=2, single_gpu=True):
    """BASELINE.json configs[3] inside bench.py: ONE `hours` h 96 kHz stereo f32 conversational stream over the `world` GPUs
    of the process group, one jt_process_audio_sharded call per rank.  Timed region (barrier + device synchronise on both
    sides, max over ranks): from the rank's window in pinned HOST memory to its owned part of the result in pinned host
    memory -- window upload, every chunk, every exchange (NCCL all-gathers of measurement blobs, region samples and halos),
    every merge, result download.  Rank 0 then runs the same stream unchunked on its one GPU (jt_process_audio_adaptive, host
    buffers) for the speed-up.  The stream is a 10 min block of the C2 recipe tiled (R = L delayed 7 samples x 0.9)."""
    import time
    import numpy as np
    from . import adapt, gpudsp, synth
    dev = torch.device("cuda", local)
    total = int(hours * 3600 * rate)
    seg = synth.stereo_from_mono(synth.podcast_like(600.0, rate, seed=12345)) if channels == 2 else synth.podcast_like(600.0, rate, seed=12345)
    seg_frames = seg.size // channels

    def window(first, n):
        """frames [first, first + n) of the tiled stream, in pinned host memory"""
        buf = torch.empty(n * channels, dtype=torch.float32, pin_memory=True)
        a = buf.numpy()
        pos = 0
        while pos < n:
            o = (first + pos) % seg_frames
            m = min(seg_frames - o, n - pos)
            a[pos * channels: (pos + m) * channels] = seg[o * channels: (o + m) * channels]
            pos += m
        return buf

    p = adapt.sharded_plan(total, rate, world, rank)
    h_in = window(p.local_first, p.n_local)
    cap = int(p.owned * 44100 / rate) + 4 * 4096 + 2 * 890820
    h_out = torch.empty(cap, dtype=torch.int16, pin_memory=True)
    comm = DistComm(dev)
    ctx.set_exchange(comm.exchange_raw, world, raw=True)
    times, last = [], None
    try:
        for rep in range(reps + 1):                    # first repetition = warm-up
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            last = adapt.process_audio_sharded_ptr(ctx, h_in.data_ptr(), p.n_local, rate, channels, gpudsp.FMT_FLT, total, world, rank,
                                                   h_out.data_ptr(), cap, False, adaptive=True)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            if rep > 0:
                times.append(float(dt[0]))
    finally:
        ctx.set_exchange(None, 1)
    first, n_out, res, an, tm = last
    keys = [k for k, _ in adapt.ShardTiming._fields_ if k not in ("reserved", "halo_bytes", "exchange_calls")]
    ph = torch.tensor([getattr(tm, k) for k in keys], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ph, op=dist.ReduceOp.MAX)
    best = min(times)
    single = None
    if single_gpu:
        del h_in
        if rank == 0:
            try:
                whole = window(0, total)
                out1 = torch.empty(int(total * 44100 / rate) + 3 * 4096, dtype=torch.int16, pin_memory=True)
                ts = []
                for rep in range(2):
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    r1, a1 = adapt.process_audio_adaptive_ptr(ctx, whole.data_ptr(), total, rate, channels, gpudsp.FMT_FLT, out1.data_ptr(), out1.numel(), False)
                    torch.cuda.synchronize()
                    ts.append(time.perf_counter() - t0)
                single = {"seconds": min(ts), "final_lufs": r1.final.input_i, "final_dbtp": r1.final.input_tp,
                          "spec_equal": bool(a1.pass2_spec == an.pass2_spec)}
                del whole, out1
            except Exception as e:          # noqa: BLE001  (host memory for an 8 GB pinned stream may not be there)
                single = {"error": str(e)}
        if world > 1:
            dist.barrier()
    if rank != 0:
        return None
    phases = {k: round(float(v), 4) for k, v in zip(keys, ph)}
    limiter = max((k for k in phases if k != "exchange"), key=lambda k: phases[k])
    out = {"workload": f"single {hours:g} h {rate} Hz {channels}-channel f32 conversational stream, one chunk per GPU per pass (BASELINE.json configs[3])",
           "entry": "jt_process_audio_sharded (one call per rank, NCCL all-gather callback)", "n_gpus": world, "seconds_per_stream": best,
           "all_seconds": times, "realtime_x": hours * 3600 / best, "frames_per_s": total / best,
           "phase_seconds_max_over_ranks": phases, "slowest_phase": limiter, "halo_bytes_rank0": int(tm.halo_bytes), "exchange_calls": int(tm.exchange_calls),
           "final_lufs": res.final.input_i, "final_dbtp": res.final.input_tp, "final_lra": res.final.input_lra, "n_out": int(res.n_out),
           "pass2_spec": an.pass2_spec.decode()}
    if single is not None:
        out["single_gpu_unchunked"] = single
        if "seconds" in single:
            out["speedup_vs_single_gpu"] = single["seconds"] / best
    return out
