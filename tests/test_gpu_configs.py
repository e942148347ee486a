"""GPU: BASELINE.json's configs at (or near) their full sizes, checked through properties the domain offers plus the oracle on
windows the oracle finishes in seconds (shapes and signals).  configs[0] lives in test_gpu_resample.py
(test_full_size_properties_config1); configs[1] at all 256 files is what bench.py runs (it asserts that the HBM-resident and
the host-buffer legs agree bit for bit) -- here its full-length files go through the host job flow against the oracle."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 2.0 ** -20
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TAPS = {0: 200, 1: 5}


def _workloads():
    import importlib.util
    p = os.path.join(ROOT, "f9-juce-resampler-studio_b200", "py", "workloads.py")
    spec = importlib.util.spec_from_file_location("f9workloads_t", p)
    m = importlib.util.module_from_spec(spec)
    sys.modules["f9workloads_t"] = m
    spec.loader.exec_module(m)
    return m


def oracle_window(O, kind, p, q, x, n0, cnt):
    """Outputs [n0, n0 + cnt) of the conversion of x from reset state, n0 a multiple of the period q: the phase pattern repeats
    every q outputs / p inputs, so the oracle restarted w periods earlier reproduces it once its memory has filled."""
    assert n0 % q == 0
    if n0 == 0:
        return O.resample_channel(kind, p / q, x, cnt)[0]
    w = (TAPS[kind] + p - 1) // p + 1
    a0 = n0 // q - w
    assert a0 >= 0
    ref, _ = O.resample_channel(kind, p / q, x[a0 * p:], cnt + w * q)
    return ref[w * q:]


# ------------------------------------------------------------------------------------------------ configs[1]
def test_config2_full_length_files_through_the_job_flow(ctx, O, f9):
    """96 kHz -> 44.1 kHz, 10 s stereo files with per-file latency, trim + tail-silence detection + WindowedSinc / Lagrange:
    latency offsets, trim points, output lengths and tail stops bit-exact, samples within 2^-20 on windows."""
    W = _workloads()
    files = [0, 1, 37, 100, 255]
    batch = W.describe("config2_256x_stereo_96k_to_44k1_trim_tail", files=1)
    caps = [W.fill_host_numpy(batch, f, 1)[0] for f in files]
    win, hop, req = int(0.1 * batch.fs_in), int(0.05 * batch.fs_in), 3
    jobs = [dict(captured=caps[i], latency_samples=W.latency_of(f) * 2, original_length=batch.src_frames, fs_in=batch.fs_in,
                 fs_out=batch.fs_out, kind=i % 2, tail=(win, hop, req, f9.TAIL_RMS, True, -90.0, 0.0)) for i, f in enumerate(files)]
    outs, _, res = ctx.process_batch(jobs)
    n_out = f9.resampled_length(batch.src_frames, batch.fs_in, batch.fs_out)
    assert n_out == 441000
    for i, f in enumerate(files):
        lat = W.latency_of(f)
        trimmed, copied = O.trim_latency(caps[i], 2 * lat, batch.src_frames)
        stop, _ = O.tail_scan(caps[i], batch.src_frames + lat, win, hop, req, 0, True, -90.0, 0.0)
        r = res[i]
        assert r["status"] == 0
        assert (r["latency_frames"], r["trim_start"], r["frames_copied"], r["out_frames"]) == (lat, lat, copied, n_out)
        assert r["tail_stop_frame"] == stop
        for n0 in (0, 147 * 1500, 147 * ((n_out - 12000) // 147)):
            ref = oracle_window(O, i % 2, 320, 147, trimmed[1], n0, 12000)
            assert np.max(np.abs(outs[i][1][n0:n0 + 12000] - ref)) <= TOL, (f, n0)


def test_job_flow_in_many_chunks_equals_one_chunk(ctx, O, f9):
    """f9_process_batch enqueues every chunk without a host wait: chunk k + 2 reuses chunk k's device arena behind it in stream
    order.  Twelve files of mixed shape (float planes and 24-bit file bytes in, float and 24-bit payload out, tail scans, DC
    removal, two ratios) cut into one-file chunks (F9_BATCH_CHUNK_MB = 1) give the results of the single-chunk call bit for bit."""
    rng = np.random.default_rng(11)
    jobs = []
    for i in range(12):
        ch = 1 + i % 2
        src = 40000 + 977 * i
        lat = 3 + 5 * i
        cap = (rng.uniform(-0.4, 0.4, (ch, src + lat + 12000)) * np.exp(-np.arange(src + lat + 12000) / 9000.0)).astype(np.float32)
        fs = [(96000, 44100), (48000, 44100), (48000, 48000)][i % 3]
        j = dict(latency_samples=lat * ch, original_length=src, fs_in=fs[0], fs_out=fs[1], kind=i % 2,
                 tail=(4800, 2400, 3, f9.TAIL_RMS, True, -60.0, 0.0), pcm24=(i % 4 != 1), remove_dc=(i % 5 == 2))
        if i % 3 == 1:
            q = np.clip(np.round(cap.astype(np.float64) * 8388608.0), -8388608, 8388607).astype(np.int32)
            inter = q.T.reshape(-1)
            raw = np.stack([(inter >> (8 * b)) & 0xff for b in range(3)], axis=-1).astype(np.uint8).reshape(-1)
            j["src_pcm"] = (raw, f9.PCM_S24LE, ch)
        else:
            j["captured"] = cap
        jobs.append(j)
    one = ctx.process_batch(jobs)
    ctx.set_option("F9_BATCH_CHUNK_MB", 1)
    try:
        many = ctx.process_batch(jobs)
        again = ctx.process_batch(jobs)                    # and once more on the arenas the call before left behind
    finally:
        ctx.clear_options()
    for got in (many, again):
        for i in range(len(jobs)):
            assert got[2][i] == one[2][i], i
            assert np.array_equal(got[0][i], one[0][i]), i
            if jobs[i]["pcm24"]:
                assert np.array_equal(got[1][i], one[1][i]), i


def test_job_flow_adjacent_payloads_go_up_as_one_copy(ctx, O, f9):
    """File payloads that follow one another in host memory (64-byte padded slots of one buffer, as a reader holding its files in
    one block has them) are uploaded run by run; latencies of every residue modulo 4 (different placements on the device: a run
    breaks where the placement the next file needs is not the one the host spacing gives) and a mono file in between.  Results equal
    those of the same payloads in separate buffers, bit for bit."""
    rng = np.random.default_rng(17)
    specs = [(2, 30000, 8), (2, 31001, 9), (2, 29003, 9), (1, 30500, 10), (2, 30007, 11), (2, 30011, 11), (2, 30013, 7), (2, 30017, 3)]
    raws, jobs_sep = [], []
    for ch, src, lat in specs:
        cap = (rng.uniform(-0.4, 0.4, (ch, src + lat + 9000)) * np.exp(-np.arange(src + lat + 9000) / 7000.0)).astype(np.float32)
        q = np.clip(np.round(cap.astype(np.float64) * 8388608.0), -8388608, 8388607).astype(np.int32)
        inter = q.T.reshape(-1)
        raws.append(np.stack([(inter >> (8 * b)) & 0xff for b in range(3)], axis=-1).astype(np.uint8).reshape(-1))
        jobs_sep.append(dict(src_pcm=(raws[-1], f9.PCM_S24LE, ch), latency_samples=lat * ch, original_length=src, fs_in=96000, fs_out=44100,
                             kind=0, tail=(4800, 2400, 3, f9.TAIL_RMS, True, -60.0, 0.0), pcm24=True))
    offs = np.concatenate([[0], np.cumsum([(r.size + 63) // 64 * 64 for r in raws])]).astype(np.int64)
    block = np.zeros(int(offs[-1]) + 64, dtype=np.uint8)
    jobs_adj = []
    for i, (r, j) in enumerate(zip(raws, jobs_sep)):
        block[offs[i]:offs[i] + r.size] = r
        k = dict(j); k["src_pcm"] = (block[offs[i]:offs[i] + r.size], f9.PCM_S24LE, specs[i][0])
        jobs_adj.append(k)
    sep = ctx.process_batch(jobs_sep)
    adj = ctx.process_batch(jobs_adj)
    for i in range(len(specs)):
        assert adj[2][i] == sep[2][i], i
        assert np.array_equal(adj[0][i], sep[0][i]), i
        assert np.array_equal(adj[1][i], sep[1][i]), i
    # and against the oracle for one file
    cap0 = O.pcm_to_planar(raws[0], f9.PCM_S24LE, 2) if hasattr(O, "pcm_to_planar") else None
    if cap0 is not None:
        trimmed, _ = O.trim_latency(cap0, 2 * specs[0][2], specs[0][1])
        ref, _ = O.resample_channel(0, 96000 / 44100, trimmed[0], adj[0][0].shape[1])
        assert np.max(np.abs(adj[0][0][0] - ref)) <= TOL


# ------------------------------------------------------------------------------------------------ configs[2]
def test_config3_64ch_10min_time_segmented(ctx, O, f9):
    """64 channels, 48 kHz -> 192 kHz, 10 minutes: 36.9 GB resident, every channel split into 8 time segments that carry their
    own halos.  Checked: segment seams and channel ends against the oracle, and the segmented result against
    the same channels converted whole."""
    torch = pytest.importorskip("torch")
    if torch.cuda.mem_get_info()[0] < 60e9:
        pytest.skip("needs 60 GB of free device memory")
    fs_in, fs_out, nch, n_in = 48000, 192000, 64, 600 * 48000
    n_out = f9.resampled_length(n_in, fs_in, fs_out)
    assert n_out == 600 * fs_out
    d_in = torch.empty((nch, n_in), dtype=torch.float32, device="cuda")
    t = torch.arange(n_in, device="cuda", dtype=torch.float64) / fs_in
    phase = 2 * np.pi * (20.0 * t + (2000.0 - 20.0) / (2 * 600.0) * t * t)           # linear sweep 20 Hz -> 2 kHz
    for c in range(nch):
        d_in[c] = (0.5 * torch.sin(phase + 0.1 * c)).to(torch.float32)
    del t, phase
    d_out = torch.empty((nch, n_out), dtype=torch.float32, device="cuda")
    seg_out = n_out // 8
    for kind in (1, 0):
        d_out.fill_(float("nan"))
        segs = []
        for c in range(nch):
            for n0 in range(0, n_out, seg_out):
                first, last = f9.segment_input_range(kind, fs_in / fs_out, n0, seg_out)
                lo, hi = max(first, 0), min(last, n_in)
                segs.append(f9.ResampleSeg(d_in[c].data_ptr() + 4 * lo, lo, hi - lo, d_out[c].data_ptr() + 4 * n0, n0, seg_out))
        arr = (f9.ResampleSeg * len(segs))(*segs)
        plan = C.c_void_p(None)
        torch.cuda.synchronize()
        assert f9.lib().f9_resample_plan_create(ctx.handle, kind, fs_in / fs_out, arr, len(segs), C.byref(plan)) == 0
        assert f9.lib().f9_resample_plan_run(plan) == 0
        ctx.synchronize()
        f9.lib().f9_plan_destroy(plan)
        assert bool(torch.isfinite(d_out[::9, ::4099]).all())
        # the same channels converted whole
        whole = torch.empty((2, n_out), dtype=torch.float32, device="cuda")
        wsegs = (f9.ResampleSeg * 2)(*[f9.ResampleSeg(d_in[c].data_ptr(), 0, n_in, whole[i].data_ptr(), 0, n_out) for i, c in enumerate((0, 63))])
        assert f9.lib().f9_resample_plan_create(ctx.handle, kind, fs_in / fs_out, wsegs, 2, C.byref(plan)) == 0
        assert f9.lib().f9_resample_plan_run(plan) == 0
        ctx.synchronize()
        f9.lib().f9_plan_destroy(plan)
        for i, c in enumerate((0, 63)):
            assert float((whole[i] - d_out[c]).abs().max()) <= TOL, (kind, c)
        # oracle at the start, across two seams and at the end of two channels (period: 1 input -> 4 outputs)
        for c in (5, 63):
            x = d_in[c].cpu().numpy()
            got = d_out[c]
            for n0 in (0, seg_out - 4000, 5 * seg_out - 4000, n_out - 8000):
                ref = oracle_window(O, kind, 1, 4, x, n0, 8000)
                assert np.max(np.abs(got[n0:n0 + 8000].cpu().numpy() - ref)) <= TOL, (kind, c, n0)
        del whole


# ------------------------------------------------------------------------------------------------ configs[3]
def test_config4_latency_detection_512_recordings(ctx, O, f9):
    """512 stereo captures of 5 s at 48 kHz: impulse 0.9 at frame d_i in [0, 65535] over noise at -80 dBFS.  findPeakPosition and
    the bounded-lag cross-correlation over +-2^16 lags must both return d_i exactly, including the exact-tie cases (two equal
    peaks keep the earliest; equal peaks in both channels keep channel 0)."""
    torch = pytest.importorskip("torch")
    n, ch, frames = 512, 2, 240000
    rng = np.random.default_rng(4)
    delays = rng.integers(0, 65536, n)
    delays[:4] = (0, 65535, 1, 65534)
    g = torch.Generator(device="cuda"); g.manual_seed(4)
    rec = torch.randn((n, ch, frames), generator=g, device="cuda", dtype=torch.float32) * (10 ** (-80 / 20))
    idx = torch.arange(n, device="cuda")
    d = torch.from_numpy(delays).cuda()
    rec[idx, 0, d] = 0.9
    rec[idx[1::2], 1, d[1::2]] = 0.9                        # odd files: equal peaks in channel 0 and channel 1
    rec[10, 0, delays[10] + 777] = 0.9                       # a second, equal, later peak
    rec[11, 1, delays[11] + 5] = -0.9                        # equal magnitude, other channel, later
    bufs = (f9.DevBuffer * n)(*[f9.DevBuffer(rec[i].data_ptr(), frames, ch, frames) for i in range(n)])
    pos = torch.full((n,), -7, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    assert f9.lib().f9_dev_find_peak_batch(ctx.handle, bufs, n, 0.1, pos.data_ptr()) == 0
    ctx.synchronize()
    assert np.array_equal(pos.cpu().numpy(), delays.astype(np.int32))
    # oracle on a few of them (scan order and tie-breaking of Source/MainComponent.cpp:950-975)
    for i in (0, 1, 10, 11, 200):
        assert O.find_peak_position(rec[i].cpu().numpy(), 0.1) == delays[i]
    # cross-correlation with the impulse stimulus (generateImpulse: 0.9 on sample 0 of a block, Source/MainComponent.cpp:934-945)
    stim = torch.zeros(256, dtype=torch.float32, device="cuda"); stim[0] = 0.9
    raw = torch.zeros(n * 24, dtype=torch.uint8, device="cuda")                      # f9_xcorr_result[n]: double, int, int, int (+ padding) = 24 bytes
    assert f9.lib().f9_dev_xcorr_peak_batch(ctx.handle, bufs, n, stim.data_ptr(), 256, -65536, 65536, raw.data_ptr()) == 0
    ctx.synchronize()
    out = np.frombuffer(raw.cpu().numpy().tobytes(), dtype=np.dtype([("value", "<f8"), ("ch", "<i4"), ("lag", "<i4"), ("pad", "<i4"), ("pad2", "<i4")]))
    assert np.array_equal(out["lag"], delays.astype(np.int32))
    assert np.all(out["ch"] == 0)
    assert np.allclose(out["value"], 0.81, rtol=0, atol=1e-3)
    # sweep stimulus on a subset: recordings = delayed sweep + noise; the correlation peak sits at the delay
    m, L = 32, 4800
    tt = np.arange(L) / 48000.0
    sweep = (0.5 * np.sin(2 * np.pi * (200.0 * tt + (8000.0 - 200.0) / (2 * tt[-1]) * tt * tt))).astype(np.float32)
    d_sw = torch.from_numpy(sweep).cuda()
    rec2 = torch.randn((m, ch, frames), generator=g, device="cuda", dtype=torch.float32) * (10 ** (-80 / 20))
    for i in range(m):
        rec2[i, i % 2, delays[i]:delays[i] + L] += d_sw
    bufs2 = (f9.DevBuffer * m)(*[f9.DevBuffer(rec2[i].data_ptr(), frames, ch, frames) for i in range(m)])
    raw2 = torch.zeros(m * 24, dtype=torch.uint8, device="cuda")
    assert f9.lib().f9_dev_xcorr_peak_batch(ctx.handle, bufs2, m, d_sw.data_ptr(), L, -65536, 65536, raw2.data_ptr()) == 0
    ctx.synchronize()
    out2 = np.frombuffer(raw2.cpu().numpy().tobytes(), dtype=out.dtype)
    assert np.array_equal(out2["lag"], delays[:m].astype(np.int32))
    assert np.array_equal(out2["ch"], np.arange(m) % 2)


def test_config4_sweep_all_512_fast_equals_exact(ctx, O, f9):
    """Config 4 with the sweep stimulus on all 512 recordings (+-2^16 lags): the candidate path (tensor-core approximation of
    every lag + exact verification of the lags that can still win) returns value, channel and lag IDENTICAL to the exact scan of
    every lag (option F9_XCORR_EXACT_ALL) -- the value bit for bit --, and to the oracle on a few.  Includes exact ties (the same
    sweep at two delays / in both channels), silent recordings and a periodic recording (candidate list overflows: exact fallback)."""
    torch = pytest.importorskip("torch")
    n, ch, frames, L = 512, 2, 240000, 4800
    rng = np.random.default_rng(44)
    delays = rng.integers(0, 65536, n)
    tt = np.arange(L) / 48000.0
    sweep = (0.5 * np.sin(2 * np.pi * (200.0 * tt + (8000.0 - 200.0) / (2 * tt[-1]) * tt * tt))).astype(np.float32)
    d_sw = torch.from_numpy(sweep).cuda()
    g = torch.Generator(device="cuda"); g.manual_seed(44)
    rec = torch.randn((n, ch, frames), generator=g, device="cuda", dtype=torch.float32) * (10 ** (-80 / 20))
    for i in range(n):
        rec[i, i % 2, delays[i]:delays[i] + L] += d_sw
    # ties: identical content at two delays (noise-free so the sums are equal bit for bit) and in both channels
    rec[20] = 0.0; rec[20, 0, 3000:3000 + L] = d_sw; rec[20, 0, 50000:50000 + L] = d_sw
    rec[21] = 0.0; rec[21, 1, 777:777 + L] = d_sw; rec[21, 0, 777:777 + L] = d_sw
    rec[22] = 0.0                                                                      # silence
    per = torch.sin(2 * np.pi * 1000.0 * torch.arange(frames, device="cuda", dtype=torch.float64) / 48000.0).to(torch.float32) * 0.3
    rec[23, 0] = per; rec[23, 1] = per                                                 # periodic: thousands of near-maxima
    bufs = (f9.DevBuffer * n)(*[f9.DevBuffer(rec[i].data_ptr(), frames, ch, frames) for i in range(n)])
    dt = np.dtype([("value", "<f8"), ("ch", "<i4"), ("lag", "<i4"), ("pad", "<i4"), ("pad2", "<i4")])

    def run(c):
        raw = torch.zeros(n * 24, dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        assert f9.lib().f9_dev_xcorr_peak_batch(c.handle, bufs, n, d_sw.data_ptr(), L, -65536, 65536, raw.data_ptr()) == 0
        c.synchronize()
        return np.frombuffer(raw.cpu().numpy().tobytes(), dtype=dt)

    fast = run(ctx)
    c2 = f9.Context(0)
    try:
        c2.set_option("F9_XCORR_EXACT_ALL", 1)
        exact = run(c2)
    finally:
        c2.close()
    assert np.array_equal(fast["lag"], exact["lag"]) and np.array_equal(fast["ch"], exact["ch"])
    assert fast["value"].tobytes() == exact["value"].tobytes()                         # the same double sums
    ok = np.ones(n, bool); ok[[20, 21, 22, 23]] = False
    assert np.array_equal(fast["lag"][ok], delays[ok].astype(np.int32)) and np.array_equal(fast["ch"][ok], (np.arange(n) % 2)[ok])
    assert (fast["lag"][20], fast["ch"][20]) == (3000, 0) and (fast["lag"][21], fast["ch"][21]) == (777, 0) and fast["ch"][22] == -1
    for i in (0, 20, 21, 23):
        found, lag, chn, val = O.xcorr_peak(rec[i].cpu().numpy(), sweep, -65536, 65536, 0.01)
        assert (lag, chn) == (int(fast["lag"][i]), int(fast["ch"][i])) and val == fast["value"][i], i


# ------------------------------------------------------------------------------------------------ configs[4]
@pytest.mark.parametrize("kind", [0, 1])
def test_config5_mixed_rates_to_48k(ctx, O, f9, kind):
    """Mixed-rate 10 s stereo files (44.1 / 48 / 88.2 / 96 / 192 kHz round-robin) to 48 kHz through the job flow, Lagrange and
    WindowedSinc: exact output lengths, oracle on windows at the start, the middle and the end of every rate."""
    rates = [44100, 48000, 88200, 96000, 192000]
    pq = {44100: (147, 160), 48000: (1, 1), 88200: (147, 80), 96000: (2, 1), 192000: (4, 1)}
    rng = np.random.default_rng(5)
    jobs, xs = [], []
    for i in range(10):
        fs = rates[i % 5]
        n = 10 * fs
        t = np.arange(n) / fs
        x = np.stack([(0.5 * np.sin(2 * np.pi * (20.0 * t + (18000.0 - 20.0) / 20.0 * t * t))).astype(np.float32),
                      rng.uniform(-0.5, 0.5, n).astype(np.float32)])
        xs.append(x)
        jobs.append(dict(captured=x, latency_samples=0, original_length=n, fs_in=fs, fs_out=48000, kind=kind))
    outs, _, res = ctx.process_batch(jobs)
    for i, x in enumerate(xs):
        fs = rates[i % 5]
        p, q = pq[fs]
        assert res[i]["status"] == 0 and res[i]["out_frames"] == 480000 and outs[i].shape == (2, 480000)
        if fs == 48000:                                        # same rate: the job flow copies (no conversion, as the reference)
            assert np.array_equal(outs[i], x)
            continue
        cnt = 6000
        for c in (0, 1):
            for n0 in (0, q * (240000 // q), q * ((480000 - cnt) // q)):
                ref = oracle_window(O, kind, p, q, x[c], n0, cnt)
                assert np.max(np.abs(outs[i][c][n0:n0 + cnt] - ref)) <= TOL, (fs, c, n0)
