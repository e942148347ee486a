#!/usr/bin/env python
"""Per-kernel rooflines of the DSP path at the configs' sizes: every device-batch entry point timed with CUDA
events on the library's stream, inputs resident in HBM and larger than L2 (or L2 flushed between repetitions), algorithmic
bytes / time against the measured HBM peak.  One JSON object per line; profiles/r01_v3_kernel_rooflines.json is this output.

    python tools/kernel_rooflines.py [--reps 10]
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

f9 = g._load_pkg()


def main():
    import torch
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=10)
    args = ap.parse_args()
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    peak = float(peaks["hbm_gbs"])
    dev = torch.device("cuda", 0)
    ctx = f9.Context(0)
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    L = f9.lib()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)          # > 126 MB of L2

    def timed(fn, reps=args.reps, do_flush=True):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(reps):
            if do_flush:
                flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream); fn(); e1.record(stream)
            e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        return ts[len(ts) // 2]

    def report(name, config, ms, alg_bytes, unit_count, unit, note=""):
        ach = alg_bytes / (ms * 1e-3) / 1e9
        print(json.dumps({"kernel": name, "config": config, "ms": round(ms, 5), "algorithmic_bytes": alg_bytes, "achieved_GBps": round(ach, 1),
                          "peak_GBps": peak, "frac": round(ach / peak, 4), "throughput": round(unit_count / (ms * 1e-3) / 1e6, 1), "unit": "M" + unit + "/s",
                          "note": note}), flush=True)

    # ---------------- config 4: 512 stereo recordings of 5 s at 48 kHz
    n, ch, frames = 512, 2, 240000
    gen = torch.Generator(device=dev); gen.manual_seed(1)
    rec = torch.randn((n, ch, frames), generator=gen, device=dev, dtype=torch.float32) * 1e-4
    d = torch.randint(0, 65536, (n,), generator=gen, device=dev)
    rec[torch.arange(n, device=dev), 0, d] = 0.9
    bufs = (f9.DevBuffer * n)(*[f9.DevBuffer(rec[i].data_ptr(), frames, ch, frames) for i in range(n)])
    pos = torch.empty(n, dtype=torch.int32, device=dev)
    ms = timed(lambda: ctx._check(L.f9_dev_find_peak_batch(ctx.handle, bufs, n, 0.1, pos.data_ptr())))
    report("peak_partial/final (findPeakPosition)", "config4: 512 x 2 x 240000", ms, 4.0 * n * ch * frames, n * ch * frames, "samples")
    sumsq = torch.empty(n, dtype=torch.float64, device=dev); pk = torch.empty(n, dtype=torch.float32, device=dev)
    ms = timed(lambda: ctx._check(L.f9_dev_stats_batch(ctx.handle, bufs, n, sumsq.data_ptr(), pk.data_ptr())))
    report("stats_partial/final (calculateRMS / noise floor)", "config4: 512 x 2 x 240000", ms, 4.0 * n * ch * frames, n * ch * frames, "samples")
    ctx.set_option("F9_RMS_TREE_SUM", 1)
    ms = timed(lambda: ctx._check(L.f9_dev_stats_batch(ctx.handle, bufs, n, sumsq.data_ptr(), pk.data_ptr())))
    report("stats_partial/final, option F9_RMS_TREE_SUM (the tree sum as it is: no order test)", "config4: 512 x 2 x 240000", ms, 4.0 * n * ch * frames, n * ch * frames, "samples")
    ctx.clear_options()
    ctx.set_option("F9_RMS_FORCE_ORDER", 1)
    ms = timed(lambda: ctx._check(L.f9_dev_stats_batch(ctx.handle, bufs, n, sumsq.data_ptr(), pk.data_ptr())), reps=3)
    report("stats_partial/final, option F9_RMS_FORCE_ORDER (EVERY buffer re-summed in the reference's order)", "config4: 512 x 2 x 240000", ms, 4.0 * n * ch * frames, n * ch * frames, "samples",
           "one CTA per buffer: the cost of one ordered re-sum is this time / ceil(512 / resident CTAs)")
    ctx.clear_options()
    ms = timed(lambda: ctx._check(L.f9_dev_latency_stats_batch(ctx.handle, bufs, n, 0.1, pos.data_ptr(), sumsq.data_ptr(), pk.data_ptr())))
    report("peak_partial<STATS>/final (findPeakPosition + calculateNoiseFloorDb, one read)", "config4: 512 x 2 x 240000", ms, 4.0 * n * ch * frames, n * ch * frames, "samples",
           "the latency measurement's two passes over a capture (MainComponent.cpp:270-279) as one")
    stim = torch.zeros(256, dtype=torch.float32, device=dev); stim[0] = 0.9
    raw = torch.zeros(n * 24, dtype=torch.uint8, device=dev)
    ms = timed(lambda: ctx._check(L.f9_dev_xcorr_peak_batch(ctx.handle, bufs, n, stim.data_ptr(), 256, -65536, 65536, raw.data_ptr())), reps=5)
    report("xc_approx + select + exact + pick, impulse stimulus (256 samples), +-2^16 lags", "config4: 512 x 2 x 240000", ms, 4.0 * n * ch * frames, n * ch * frames, "samples",
           "candidates on the tensor cores, exact verification")
    tt = np.arange(4800) / 48000.0
    sweep = torch.from_numpy((0.5 * np.sin(2 * np.pi * (200.0 * tt + 7800.0 / (2 * tt[-1]) * tt * tt))).astype(np.float32)).to(dev)
    # recordings of the sweep measurement: the delayed sweep over the same noise (the impulse removed)
    rec[torch.arange(n, device=dev), 0, d] = 0.0
    for i in range(n):
        rec[i, i % 2, int(d[i]):int(d[i]) + 4800] += sweep
    ms = timed(lambda: ctx._check(L.f9_dev_xcorr_peak_batch(ctx.handle, bufs, n, sweep.data_ptr(), 4800, -65536, 65536, raw.data_ptr())), reps=5, do_flush=False)
    out = np.frombuffer(raw.cpu().numpy().tobytes(), dtype=np.dtype([("value", "<f8"), ("ch", "<i4"), ("lag", "<i4"), ("pad", "<i4"), ("pad2", "<i4")]))
    assert np.array_equal(out["lag"], d.cpu().numpy().astype(np.int32))
    report("xc_approx (mma.sync fp16, every lag) + xc_select + xc_exact (FP64 chains of the candidates) + xc_pick, sweep stimulus (4800 samples), +-2^16 lags",
           "config4: 512 x 2 x 240000", ms, 4.0 * n * ch * frames, n * ch * frames, "samples",
           "%.1f TFLOP/s of fp16 tensor work (2 x 131073 lags x 4800 taps per channel); argmax, channel and value identical to the exact scan" % (2.0 * n * ch * 131073 * 4800 / (ms * 1e-3) / 1e12))
    m = 64
    cx = f9.Context(0); cx.set_stream(stream.cuda_stream); cx.set_option("F9_XCORR_EXACT_ALL", 1)
    bufs_s = (f9.DevBuffer * m)(*[f9.DevBuffer(rec[i].data_ptr(), frames, ch, frames) for i in range(m)])
    ms = timed(lambda: cx._check(L.f9_dev_xcorr_peak_batch(cx.handle, bufs_s, m, sweep.data_ptr(), 4800, -65536, 65536, raw.data_ptr())), reps=3, do_flush=False)
    report("xcorr_partial/final (exact FP64 sums of every lag: option F9_XCORR_EXACT_ALL, the fallback path), sweep stimulus (4800 samples), +-2^16 lags", "config4 subset: 64 x 2 x 240000", ms, 4.0 * m * ch * frames, m * ch * frames, "samples",
           "FP64-bound: %.2f TFLOP/s (FP64 FMA); 512 recordings = 8x this" % (2.0 * m * ch * 131073 * 4800 / (ms * 1e-3) / 1e12))
    cx.synchronize(); cx.close()
    del rec

    # ---------------- config 2: trim + tail scan over 256 stereo captures (96 kHz, 10 s + latency + 0.5 s)
    nf, src, cap = 256, 960000, 1040704
    caps = torch.randn((nf, 2, cap), generator=gen, device=dev, dtype=torch.float32) * 1e-5
    outs = torch.empty((nf, 2, src), dtype=torch.float32, device=dev)
    cb = (f9.DevBuffer * nf)(*[f9.DevBuffer(caps[i].data_ptr(), cap, 2, cap) for i in range(nf)])
    ob = (f9.DevBuffer * nf)(*[f9.DevBuffer(outs[i].data_ptr(), src, 2, src) for i in range(nf)])
    lat = (C.c_int * nf)(*[2 * (128 * i + 7) for i in range(nf)])
    ms = timed(lambda: ctx._check(L.f9_dev_trim_batch(ctx.handle, cb, lat, ob, nf, 0)), do_flush=False)
    report("trim_kernel (trimLatency as a copy)", "config2: 256 x 2 x 960000", ms, 8.0 * nf * 2 * src, nf * 2 * src, "samples", "in the job flow the trim is a pointer offset fused into the resampler")
    ms = timed(lambda: ctx._check(L.f9_dev_trim_batch(ctx.handle, cb, lat, ob, nf, 1)), do_flush=False)
    report("dc_sum_src + trim_kernel<SUB> (trimLatency + removeDCOffset, fused)", "config2: 256 x 2 x 960000", ms, (4.0 + 8.0) * nf * 2 * src, nf * 2 * src, "samples",
           "mean of the copied region (one read), then read + subtract + write; the unfused form moved 20 bytes per sample")
    win, hop = 9600, 4800
    polls = (cap - src) // hop
    tails = (f9.TailParams * nf)(*[f9.TailParams(src + 128 * i + 7, win, hop, 3, f9.TAIL_RMS, 1, -90.0, 0.0) for i in range(nf)])
    stops = torch.empty(nf, dtype=torch.int64, device=dev); flags = torch.empty(nf * polls, dtype=torch.int32, device=dev)
    ms = timed(lambda: ctx._check(L.f9_dev_tail_scan_batch(ctx.handle, cb, tails, nf, stops.data_ptr(), flags.data_ptr(), polls)))
    tail_samples = sum(2 * min(polls * hop + win, cap - (src + 128 * i + 7)) for i in range(nf))
    report("tail_window/tail_runs (reverb-tail silence scan)", "config2: 256 captures, 100 ms window / 50 ms hop", ms, 4.0 * tail_samples, tail_samples, "samples",
           "windows overlap 2x: the second read of a sample is an L2 hit")
    # 24-bit payload both ways on the converted outputs' size
    n_out = 441000
    planar = torch.randn((nf * 2, n_out), generator=gen, device=dev, dtype=torch.float32) * 0.1
    pcm = torch.empty(nf * 2 * n_out * 3, dtype=torch.uint8, device=dev)
    pbufs = (f9.DevBuffer * nf)(*[f9.DevBuffer(planar[2 * i].data_ptr(), n_out, 2, n_out) for i in range(nf)])
    pptrs = (C.c_void_p * nf)(*[pcm.data_ptr() + 6 * i * n_out for i in range(nf)])
    ms = timed(lambda: ctx._check(L.f9_dev_planar_to_pcm24_batch(ctx.handle, pbufs, pptrs, nf)), reps=5, do_flush=False)
    report("planar_to_pcm24_batch_kernel (24-bit WAV payload), one launch", "config2 outputs: 256 x 2 x 441000", ms, 7.0 * nf * 2 * n_out, nf * 2 * n_out, "samples",
           "batched entry point (the per-file form was launch-bound: 1.65 ms, 14.6 %)")
    back = torch.empty((nf * 2, n_out), dtype=torch.float32, device=dev)
    bbufs = (f9.DevBuffer * nf)(*[f9.DevBuffer(back[2 * i].data_ptr(), n_out, 2, n_out) for i in range(nf)])
    ms = timed(lambda: ctx._check(L.f9_dev_pcm_to_planar_batch(ctx.handle, pptrs, f9.PCM_S24LE, 2, bbufs, nf)), reps=5, do_flush=False)
    report("pcm_to_planar_batch_kernel (24-bit -> float planes), one launch", "config2 outputs: 256 x 2 x 441000", ms, 7.0 * nf * 2 * n_out, nf * 2 * n_out, "samples",
           "batched entry point (the per-file form was launch-bound: 1.59 ms, 15.2 %)")
    assert bool(((back - planar).abs() <= 2.0 ** -23).all())             # 24-bit round trip of |x| < 1
    del caps, outs, planar, pcm, back

    # ---------------- resampler on the other ratios (config 5's rates, config 3's shape) and juce::ResamplingAudioSource
    def plan_time(kind, fs_in, fs_out, nch, n_in, label):
        x = torch.randn((nch, n_in), generator=gen, device=dev, dtype=torch.float32) * 0.25
        no = f9.resampled_length(n_in, fs_in, fs_out)
        y = torch.empty((nch, no), dtype=torch.float32, device=dev)
        segs = (f9.ResampleSeg * nch)(*[f9.ResampleSeg(x[c].data_ptr(), 0, n_in, y[c].data_ptr(), 0, no) for c in range(nch)])
        plan = C.c_void_p(None)
        ctx._check(L.f9_resample_plan_create(ctx.handle, kind, fs_in / fs_out, segs, nch, C.byref(plan)))
        ms = timed(lambda: ctx._check(L.f9_resample_plan_run(plan)), reps=5, do_flush=nch * n_in * 4 < (256 << 20))
        L.f9_plan_destroy(plan)
        name = ("hankel_fir_kernel WindowedSinc " if fs_out % fs_in == 0 and fs_out > fs_in else "umma_fir_kernel WindowedSinc ") if kind == 0 else \
               ("umma_fir_kernel Lagrange " if fs_in % fs_out == 0 and fs_in > fs_out else "short_kernel Lagrange ")
        report(name + "%d -> %d" % (fs_in, fs_out), label, ms,
               4.0 * nch * (n_in + no), nch * no, "samples")
        return x, y, no
    for kind in (0, 1):
        for fs in (44100, 88200, 96000, 192000):
            plan_time(kind, fs, 48000, 512, 10 * fs, "config5 rate: 256 stereo files of 10 s")
        plan_time(kind, 44100, 48000, 2, 60 * 44100, "config1: one 60 s stereo file (launch-latency bound)")
    for kind in (0, 1):
        plan_time(kind, 48000, 192000, 64, 120 * 48000, "config3 shape: 64 channels, 2 of the 10 minutes")
    plan_time(0, 48000, 96000, 512, 10 * 48000, "1:2 upsampling, 256 stereo files of 10 s")
    for fs_in, fs_out in ((96000, 44100), (48000, 192000)):
        nch, n_in = 512, 10 * fs_in
        ratio = fs_in / fs_out
        no = int(n_in / ratio) - 8
        x = torch.randn((nch, n_in), generator=gen, device=dev, dtype=torch.float32) * 0.25
        y = torch.empty((nch, no), dtype=torch.float32, device=dev)
        sf = L.f9_ras_scratch_frames(ratio, no)
        scr = torch.empty((nch, sf), dtype=torch.float32, device=dev)
        ms = timed(lambda: ctx._check(L.f9_dev_ras_convert(ctx.handle, x.data_ptr(), n_in, nch, n_in, ratio, y.data_ptr(), no, no, scr.data_ptr(), sf)), reps=5, do_flush=False)
        report("ras_biquad + ras_lerp (juce::ResamplingAudioSource) %d -> %d" % (fs_in, fs_out), "512 channels of 10 s", ms, 4.0 * nch * (n_in + no), nch * no, "samples",
               "chunk-parallel FP64 biquad: latency of the double-precision recurrence, not bandwidth")
        del x, y, scr
    ctx.close()


if __name__ == "__main__":
    main()
