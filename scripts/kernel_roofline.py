"""Per-kernel roofline table from ONE `ncu --set full` pass over every launch of a 60 min ProcessAudio step
(scripts/gpu_profile_all.sh -> gpurun_out/raw_<tag>_all.csv): for every kernel (function + template arguments) the launch
with the longest duration -- the whole-stream launch -- gives duration, DRAM bytes, achieved GB/s on the kernel's ALGORITHMIC
bytes, the fraction of the measured HBM peak, the traffic ratio (DRAM bytes / algorithmic bytes) and the resource that
actually limits it.  Writes profiles/kernel_roofline_<round>.json (read by bench.py for `roofline.traffic`) and a .md table.

usage: python scripts/kernel_roofline.py r2j [r2]"""
import collections
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
rnd = sys.argv[2] if len(sys.argv) > 2 else "r2"
N = 60 * 60 * 48000                 # input samples of the profiled stream (48 kHz mono)
M = N * 44100 // 48000              # samples at the output rate

# kernel -> (bench.py timing group, algorithmic bytes of the whole-stream launch, what those bytes are)
ALG = {
    "k_anlmdn": ("anlmdn", 8 * N, "f32 in + f32 out (since the screen: only the listed hops)"),
    "k_nlm_screen<3>": ("anlmdn:screen", 8 * N, "f32 in + f32 out (pass-through of every hop the screen clears)"),
    "k_dc_interp": ("adeclick:interp", 16 * M, "f64 in + f64 out at 44.1 kHz"),
    "k_dc_detect": ("adeclick:detect", 8 * M, "f64 in (flags out are tiny)"),
    "k_dc_autocorr": ("adeclick:autocorr", 8 * M, "f64 in (25 lags per window out)"),
    "k_dc_levinson": ("adeclick:levinson", 0, "per-window AR model only"),
    "k_envelope": ("envelope_follower", 16 * N, "f64 in + f64 envelope out"),
    "k_gate_apply": ("agate_gain", 24 * N, "f64 in + f64 envelope in + f64 out"),
    "k_comp_apply": ("acompressor_gain", 24 * N, "f64 in + f64 envelope in + f64 out"),
    "k_alimiter": ("alimiter", 16 * M, "f64 in + f64 out at 44.1 kHz"),
    "k_afftdn_fwd": ("afftdn:fwd", 4 * N + 8 * N * 1025 // 1024, "f32 in + complex f32 spectra out (2x overlapped hops, 1025 bins)"),
    "k_afftdn_gain": ("afftdn:gain", 2 * 8 * N * 1025 // 1024, "spectra in + spectra out"),
    "k_afftdn_synth": ("afftdn:synth", 8 * N * 1025 // 1024 + 8 * N, "spectra in + overlapped f32 frames out"),
    "k_afftdn_ola": ("afftdn:ola", 8 * N + 4 * N, "overlapped frames in + f32 out"),
    "k_afftdn_bandsum": ("afftdn:bands", 8 * N * 1025 // 1024, "spectra in"),
    "k_afftdn_bandrec": ("afftdn:bands", 0, "15 bands per hop only"),
    "k_afftdn_floor": ("afftdn:floor", 0, "per-hop scalars only"),
    "k_spectral": ("aspectralstats", 4 * N, "f32 in (13 statistics per shown hop out)"),
    "k_spectral_flux": ("aspectralstats", 0, "magnitude rows of the shown hops"),
    "k_r128_ticks<float, 0>": ("r128_kweight_ticks", 4 * N, "f32 in (per-tick energies out)"),
    "k_kw_blocks<float, 0, 0, 128>": ("r128_kweight_ticks", 4 * N, "f32 in (forced end state per 96-sample block out)"),
    "k_kw_blocks<float, 0, 1, 128>": ("r128_kweight_ticks", 4 * N, "f32 in (block energies / peaks out)"),
    "k_kw_blocks<double, 0, 0, 64>": ("r128_kweight_ticks", 8 * M, "f64 in at 44.1 kHz"),
    "k_kw_blocks<double, 0, 1, 64>": ("r128_kweight_ticks", 8 * M, "f64 in at 44.1 kHz"),
    "k_kw_blocks<short, 0, 0, 128>": ("r128_kweight_ticks", 0, "regions only"),
    "k_kw_blocks<short, 0, 1, 128>": ("r128_kweight_ticks", 0, "regions only"),
    "k_kw_scan": ("r128_kweight_ticks", 0, "4 doubles per block"), "k_kw_fold": ("r128_kweight_ticks", 0, "2 doubles per block"),
    "k_envelope_tiles<1>": ("envelope_follower", 16 * N, "f64 in + f64 envelope out"),
    "k_envelope_tiles<0>": ("envelope_follower", 16 * N, "f64 in + f64 envelope out"),
    "k_biquad_tiles<true>": ("biquad", 8 * N, "f32 in + f32 out"), "k_biquad_tiles<false>": ("biquad", 8 * N, "f32 in + f32 out"),
    "k_copy_small": ("copy_small", 0, "per-tick values / statistics rows to pinned host memory"),
    "k_fd_scan<false>": ("flac_decode:scan", 0.86 * M, "FLAC stream in (~0.86 B / sample)"), "k_fd_scan<true>": ("flac_decode:scan", 0.86 * M, "FLAC stream in"),
    "k_fd_link": ("flac_decode:link", 0.86 * M, "FLAC stream in (CRC-16)"), "k_fd_frames": ("flac_decode:frames", 0.86 * M + 4 * M, "stream in + planar int32 out"),
    "k_fd_output<short>": ("flac_decode:output", 6 * M, "planar int32 in + s16 out"),
    "k_r128_ticks<float, 1>": ("r128_kweight_ticks:192k", 4 * 4 * M, "f32 in at 192 kHz"),
    "k_r128_ticks<double, 1>": ("r128_kweight_ticks:192k", 8 * M, "f64 in at 44.1 kHz (loudnorm's output meter)"),
    "k_swr_small<float, 4, 1>": ("truepeak_oversample:small_f64", 4 * N, "f32 in at 48 kHz (per-tick maxima out)"),
    "k_swr_phase_f64<float, 32, 1, 640>": ("truepeak_oversample:phase_f64", 4 * M, "f32 in at 44.1 kHz (per-tick maxima out)"),
    "k_swr_phase_f64<float, 36, 0, 160>": ("swr_resample:phase_f64", 4 * N + 8 * M, "f32 in at 48 kHz + f64 out at 44.1 kHz"),
    "k_swr_slot_f32<short, float, 32, 5, 2>": ("swr_resample:slot_f32_up", 2 * M + 4 * 4 * M, "s16 in at 44.1 kHz + f32 out at 192 kHz"),
    "k_biquad<float, float, 1>": ("biquad", 8 * N, "f32 in + f32 out"),
    "k_band_rms<float, float>": ("band_rms", 0, "elected region only (<= 60 s)"),
    "k_astats_a<double>": ("astats:sums_hist", 8 * N, "f64 in"), "k_astats_a<float>": ("astats:sums_hist", 4 * N, "f32 in"),
    "k_astats_b<double>": ("astats:extrema_runs", 8 * N, "f64 in"), "k_astats_b<float>": ("astats:extrema_runs", 4 * N, "f32 in"),
    "k_astats_c1<double>": ("astats:rms_scan", 8 * N, "f64 in"), "k_astats_c1<float>": ("astats:rms_scan", 4 * N, "f32 in"),
    "k_astats_c2<double>": ("astats:rms_scan", 8 * N, "f64 in"), "k_astats_c2<float>": ("astats:rms_scan", 4 * N, "f32 in"),
    "k_astats_nf<double>": ("astats:noise_floor", 8 * N, "f64 in"), "k_astats_nf<float>": ("astats:noise_floor", 4 * N, "f32 in"),
    "k_convert<float, double>": ("convert", 12 * N, "f32 in + f64 out"), "k_convert<double, float>": ("convert", 12 * N, "f64 in + f32 out"),
    "k_convert<double, short>": ("convert", 10 * M, "f64 in + s16 out"), "k_convert<short, double>": ("convert", 10 * M, "s16 in + f64 out"),
    "k_scale_f64": ("loudnorm_linear_gain", 16 * M, "f64 in + f64 out"),
    "k_raw_frame_stats<float>": ("raw_frame_stats", 4 * N, "f32 in"),
    "k_flac_frames": ("flac:frames", 2 * M, "s16 in (+ ~0.9 B/sample of frames out)"),
    "k_flac_pack": ("flac:pack", 0, "frames in + stream out (~1 B/sample each)"),
}
STALLS = ["long_scoreboard", "short_scoreboard", "wait", "barrier", "math_pipe_throttle", "mio_throttle", "lg_throttle", "dispatch_stall",
          "branch_resolving", "no_instruction", "not_selected", "tex_throttle", "membar", "drain", "sleeping"]
STALL_MEANS = {"long_scoreboard": "waiting on global / local memory", "short_scoreboard": "waiting on shared memory / special-function results",
               "wait": "fixed-latency dependent issue (a serial arithmetic chain)", "barrier": "CTA barriers", "math_pipe_throttle": "a math pipe is saturated",
               "mio_throttle": "shared-memory / special-function instruction queue full", "lg_throttle": "global-memory instruction queue full",
               "not_selected": "enough eligible warps: issue-slot bound", "branch_resolving": "branch resolution", "no_instruction": "instruction fetch",
               "dispatch_stall": "dispatch", "tex_throttle": "texture queue", "membar": "memory barrier", "drain": "store drain", "sleeping": "sleeping"}


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "MEASURED_PEAKS.json"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def main():
    src = os.path.join(ROOT, "gpurun_out", f"raw_{tag}_all.csv")
    rows = list(csv.reader(l for l in open(src) if not l.startswith("==")))
    hdr, units, rows = rows[0], rows[1], rows[2:]
    ix = {k: i for i, k in enumerate(hdr)}

    def val(row, key, default=float("nan")):
        try:
            return float(row[ix[key]].replace(",", ""))
        except Exception:
            return default

    def to_unit(row, key, want):
        v, u = val(row, key), units[ix[key]]
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3} if want == "ms" else {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
        return v * scale.get(u, 1.0)

    pk, pk_src = peak()
    by = collections.OrderedDict()
    for row in rows:
        name = row[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("<unnamed>::", "").strip()
        ms = to_unit(row, "gpu__time_duration.sum", "ms")
        e = by.setdefault(name, {"launches": 0, "total_ms": 0.0, "row": None, "ms": -1.0})
        e["launches"] += 1
        e["total_ms"] += ms
        if ms > e["ms"]:
            e["ms"], e["row"] = ms, row
    out = {}
    for name, e in by.items():
        row = e["row"]
        group, alg, what = ALG.get(name, (None, 0, "n/a"))
        dram = to_unit(row, "dram__bytes_read.sum", "B") + to_unit(row, "dram__bytes_write.sum", "B")
        stalls = {s: val(row, f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio", 0.0) for s in STALLS}
        top = max(stalls, key=lambda s: stalls[s])
        issue = val(row, "smsp__issue_active.avg.pct_of_peak_sustained_active")
        fp64 = val(row, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active")
        dram_pct = val(row, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")
        if dram_pct >= 60:
            limiter = f"HBM ({dram_pct:.0f} % of peak DRAM throughput)"
        elif issue >= 60:
            limiter = f"instruction issue ({issue:.0f} % issue-active)"
        elif fp64 >= 60:
            limiter = f"FP64 pipe ({fp64:.0f} % active)"
        else:
            limiter = f"latency: {top} {stalls[top]:.1f}/issue -- {STALL_MEANS.get(top, top)}"
        ach = alg / (e["ms"] * 1e-3) / 1e9 if alg and e["ms"] > 0 else None
        out[name] = {
            "group": group, "launches_per_step": e["launches"], "total_ms_per_step": round(e["total_ms"], 3), "ms": round(e["ms"], 4),
            "dram_bytes_per_launch": dram, "algorithmic_bytes_per_launch": alg, "algorithmic_bytes_are": what,
            "achieved_GBps": None if ach is None else round(ach, 1), "frac_of_hbm_peak": None if ach is None else round(ach / pk, 4),
            "traffic_ratio": None if not alg else round(dram / alg, 2), "limiter": limiter,
            "issue_active_pct": round(issue, 1), "warps_active_pct": round(val(row, "sm__warps_active.avg.pct_of_peak_sustained_active"), 1),
            "fp64_pipe_pct": round(fp64, 1), "dram_throughput_pct": round(dram_pct, 1),
            "l1_hit_pct": round(val(row, "l1tex__t_sector_hit_rate.pct"), 1), "l2_hit_pct": round(val(row, "lts__t_sector_hit_rate.pct"), 1),
            "registers": int(val(row, "launch__registers_per_thread", 0)), "grid": int(val(row, "launch__grid_size", 0)), "block": int(val(row, "launch__block_size", 0)),
            "occupancy_limit": {k: int(val(row, f"launch__occupancy_limit_{k}", 0)) for k in ("registers", "shared_mem", "warps", "blocks")},
            "top_stalls_per_issue": {s: round(v, 2) for s, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:3]},
        }
    # bench.py looks `traffic` up by timing group: the group's heaviest kernel
    groups = {}
    for name, k in out.items():
        g = k["group"]
        if g and (g not in groups or k["ms"] > out[groups[g]]["ms"]):
            groups[g] = name
    doc = {"source": f"gpurun_out/raw_{tag}_all.csv (ncu --set full --clock-control none, every launch of one 60 min adaptive ProcessAudio step + FLAC encode; scripts/gpu_profile_all.sh)",
           "hbm_peak_GBps": pk, "hbm_peak_source": pk_src, "samples_in": N, "samples_out": M,
           "kernels": {**out, **{g: dict(out[n], kernel=n) for g, n in groups.items() if g not in out}}, "by_function": list(out)}
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    json.dump(doc, open(os.path.join(ROOT, "profiles", f"kernel_roofline_{rnd}.json"), "w"), indent=1)
    md = [f"# Per-kernel roofline, round {rnd[1:]} ({tag})", "",
          "`ncu --set full --clock-control none -k regex:k_` over every launch of ONE 60 min 48 kHz mono ProcessAudio (adaptive) step plus the FLAC",
          f"encode of its result (`scripts/gpu_profile_all.sh {tag} 60`, B200).  One row per kernel: its longest launch (the whole-stream one).",
          f"achieved = algorithmic bytes / duration; frac = achieved / {pk:.1f} GB/s ({pk_src}); traffic = DRAM bytes read + written / algorithmic bytes.",
          "ncu durations are cold-cache and serialised; shares, not absolutes, compare with `bench.py`'s `kernels_ms_per_step`.", "",
          "| kernel | launches | ms (longest) | ms (all) | DRAM GB | algorithmic GB | achieved GB/s | frac | traffic | issue % | warps % | fp64 % | limiter |",
          "|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---|"]
    for name, k in sorted(out.items(), key=lambda kv: -kv[1]["total_ms_per_step"]):
        md.append(f"| `{name}` | {k['launches_per_step']} | {k['ms']:.3f} | {k['total_ms_per_step']:.3f} | {k['dram_bytes_per_launch'] / 1e9:.3f} | "
                  f"{k['algorithmic_bytes_per_launch'] / 1e9:.3f} | {'' if k['achieved_GBps'] is None else k['achieved_GBps']} | "
                  f"{'' if k['frac_of_hbm_peak'] is None else k['frac_of_hbm_peak']} | {'' if k['traffic_ratio'] is None else k['traffic_ratio']} | "
                  f"{k['issue_active_pct']} | {k['warps_active_pct']} | {k['fp64_pipe_pct']} | {k['limiter']} |")
    tot = sum(k["total_ms_per_step"] for k in out.values())
    md += ["", f"total kernel time under ncu: {tot:.1f} ms over {sum(k['launches_per_step'] for k in out.values())} launches", ""]
    open(os.path.join(ROOT, "profiles", f"kernel_roofline_{rnd}.md"), "w").write("\n".join(md) + "\n")
    print("\n".join(md))


if __name__ == "__main__":
    main()
