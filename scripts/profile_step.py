"""One ProcessAudio (adaptive) step over the bench's 60 min stream, device resident -- the process ncu wraps
(scripts/gpu_profile_all.sh).  Never a bench value."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--minutes", type=int, default=60)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--flac", action="store_true", help="also encode the result as FLAC")
    a = ap.parse_args()
    import torch
    from jivetalking_b200 import adapt, gpudsp
    x = bench.make_input(12345, a.minutes)
    n = len(x)
    d_in = torch.from_numpy(x).cuda()
    out_cap = int(n * 44100 / bench.RATE) + 3 * 4096
    d_out = torch.empty(out_cap, dtype=torch.int16, device="cuda")
    with gpudsp.Context(0) as ctx:
        for _ in range(a.steps):
            res, an = adapt.process_audio_adaptive_ptr(ctx, d_in.data_ptr(), n, bench.RATE, 1, gpudsp.FMT_FLT, d_out.data_ptr(), out_cap, True)
        if a.flac:
            cap_b = int(gpudsp.lib().jt_flac_max_bytes(int(res.n_out), 4096))
            d_flac = torch.empty(cap_b, dtype=torch.uint8, device="cuda")
            nb = ctx.flac_encode_ptr(d_out.data_ptr(), int(res.n_out), 44100, 4096, d_flac.data_ptr(), cap_b, True)
            d_back = torch.empty(int(res.n_out) + 4096, dtype=torch.int16, device="cuda")
            ctx.decode_ptr("flac", d_flac.data_ptr(), nb, d_back.data_ptr(), int(res.n_out) + 4096, True)
        torch.cuda.synchronize()
    print("profile_step: n_out", int(res.n_out), "I", res.final.input_i)


if __name__ == "__main__":
    main()
