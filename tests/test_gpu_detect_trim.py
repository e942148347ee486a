"""GPU parity (through the C ABI) for latency detection, tail silence, trimming and format conversion.

Integer / index / byte results are compared bit for bit with the oracle; float scalars (RMS, noise floor)
to one float ulp; DC removal to the stated tolerance (the reference sums sequentially in float).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SIZES = [(1, 1), (2, 7), (2, 255), (2, 4096), (3, 16385), (2, 220500), (5, 70001)]


def rnd(shape, seed, scale=0.5):
    return np.random.default_rng(seed).uniform(-scale, scale, shape).astype(np.float32)


# ---------------------------------------------------------------- findPeakPosition
@pytest.mark.parametrize("shape", SIZES)
def test_find_peak_random(ctx, O, shape):
    x = rnd(shape, sum(shape))
    for thr in (0.1, 0.49999, 2.0):
        assert ctx.find_peak_position(x, thr) == O.find_peak_position(x, thr)


def test_find_peak_ties_zero_nan(ctx, O):
    x = np.zeros((2, 50000), np.float32)
    assert ctx.find_peak_position(x, 0.1) == -1
    x[1, 100] = 0.5; x[0, 40000] = 0.5                      # equal peaks in ch0 and ch1, far apart in chunks
    assert ctx.find_peak_position(x, 0.1) == O.find_peak_position(x, 0.1) == 40000
    x[0, 41000] = -0.5; x[0, 16384] = 0.5                   # more ties, chunk boundary
    assert ctx.find_peak_position(x, 0.1) == O.find_peak_position(x, 0.1) == 16384
    x[0, 3] = np.nan; x[1, 77] = np.inf
    assert ctx.find_peak_position(x, 0.1) == O.find_peak_position(x, 0.1) == 77
    assert ctx.find_peak_position(np.zeros((2, 0), np.float32), 0.1) == -1


def test_find_peak_every_alignment(ctx, O):
    base = rnd((1, 4200), 11)
    for off in range(0, 9):
        for n in (1, 2, 3, 4, 5, 63, 64, 65, 1023, 4100):
            x = np.ascontiguousarray(base[:, off:off + n])
            assert ctx.find_peak_position(x, 0.0) == O.find_peak_position(x, 0.0)


def test_find_peak_interleaved(ctx, O):
    a = rnd(100001, 5)
    assert ctx.find_peak_interleaved(a, 0.1) == O.find_peak_interleaved(a, 0.1)
    assert ctx.find_peak_interleaved(a, 0.6) == O.find_peak_interleaved(a, 0.6)
    z = np.zeros(1000, np.float32)
    assert ctx.find_peak_interleaved(z, 0.1) == O.find_peak_interleaved(z, 0.1) == (0, False)


def test_latency_measurement_flow(ctx, O):
    """Source/MainComponent.cpp:265-294: impulse capture -> peak -> *2 -> noise floor."""
    fs = 48000
    rng = np.random.default_rng(0)
    for d in (0, 7, 512, 65535):
        cap = (rng.standard_normal((2, 5 * fs)) * 1e-4).astype(np.float32)      # -80 dBFS noise
        cap[:, d] += 0.9
        pk = ctx.find_peak_position(cap, 0.1)
        assert pk == O.find_peak_position(cap, 0.1) == d
        nf_gpu, nf_cpu = ctx.calculate_noise_floor_db(cap), O.noise_floor_db(cap)
        assert nf_gpu.tobytes() == nf_cpu.tobytes()                              # calculateNoiseFloorDb bit for bit


# ---------------------------------------------------------------- RMS
@pytest.mark.parametrize("shape", SIZES + [(2, 240000), (2, 960000)])
def test_rms_bit_exact(ctx, O, shape):
    """calculateRMS (Source/MainComponent.cpp:983-1004) returns the reference's float bit for bit: the tree sum is redone in
    the reference's sequential order whenever the order could change the rounded result (f9_scan.cu: rms_order_dependent)."""
    for seed, scale in ((3, 0.5), (4, 1.0), (5, 1e-4)):
        x = rnd(shape, seed + sum(shape), scale)
        g, c = ctx.calculate_rms(x), O.calculate_rms(x)
        assert g.tobytes() == c.tobytes(), (seed, scale)
        assert ctx.calculate_noise_floor_db(x).tobytes() == O.noise_floor_db(x).tobytes()
    assert float(ctx.calculate_rms(np.zeros((2, 0), np.float32))) == 0.0
    assert float(ctx.calculate_rms(np.zeros(shape, np.float32))) == 0.0


def test_rms_reference_order_path(ctx, O, f9):
    """The sequential re-sum itself (forced through the F9_RMS_FORCE_ORDER option) gives the reference's double sum, so the
    float result equals the oracle's for any data; many small buffers also sweep the automatic trigger."""
    c2 = f9.Context(0)
    try:
        c2.set_option("F9_RMS_FORCE_ORDER", 1)
        for shape in [(1, 1), (2, 255), (3, 16385), (2, 220500)]:
            x = rnd(shape, 11 + sum(shape), 0.7)
            assert c2.calculate_rms(x).tobytes() == O.calculate_rms(x).tobytes()
            assert c2.calculate_noise_floor_db(x).tobytes() == O.noise_floor_db(x).tobytes()
    finally:
        c2.close()
    rng = np.random.default_rng(12)
    for trial in range(300):
        n = int(rng.integers(1, 700))
        x = rng.uniform(-1, 1, (2, n)).astype(np.float32)
        assert ctx.calculate_rms(x).tobytes() == O.calculate_rms(x).tobytes(), trial


# ---------------------------------------------------------------- tail predicates and scan
def test_tail_predicates_match(ctx, O):
    rng = np.random.default_rng(2)
    for trial in range(40):
        db = rng.uniform(-130, -60)
        w = (rng.standard_normal((2, 2048)) * 10 ** (db / 20)).astype(np.float32)
        for has_nf, nf, mg in ((True, -96.0, 10.0), (True, -98.0, 15.0), (False, 0.0, 10.0), (True, -70.0, 0.0)):
            assert ctx.is_reverb_tail_below_noise_floor(w, has_nf, nf, mg) == O.tail_below_floor(w, has_nf, nf, mg)
            iw = O.interleave(w)
            assert ctx.is_reverb_tail_below_noise_floor_swift(iw, has_nf, nf, mg) == O.tail_below_floor_swift(iw, has_nf, nf, mg)
    z = np.zeros((2, 2048), np.float32)
    assert ctx.is_reverb_tail_below_noise_floor(z, True, -96.0, 10.0) == O.tail_below_floor(z, True, -96.0, 10.0)
    assert ctx.is_reverb_tail_below_noise_floor_swift(z.ravel(), True, -96.0, 10.0) == O.tail_below_floor_swift(z.ravel(), True, -96.0, 10.0)
    assert ctx.is_reverb_tail_below_noise_floor_swift(z.ravel(), True, -200.0, 0.0) == O.tail_below_floor_swift(z.ravel(), True, -200.0, 0.0)


def test_tail_predicate_on_the_threshold(ctx, O):
    """Windows whose RMS sits within an ulp of the decision boundary: the guard band must reproduce the
    sequential reference decision exactly."""
    rng = np.random.default_rng(4)
    nf, mg = -96.0, 10.0
    thr_db = float(O.noise_floor_threshold_db(True, nf, mg))
    target = 10 ** (thr_db / 20)
    mismatches = 0
    for trial in range(200):
        w = rng.standard_normal((2, 2048)).astype(np.float32)
        w *= np.float32(target / np.sqrt(np.mean(w.astype(np.float64) ** 2)))
        w *= np.float32(1.0 + rng.uniform(-3e-7, 3e-7))
        mismatches += ctx.is_reverb_tail_below_noise_floor(w, True, nf, mg) != O.tail_below_floor(w, True, nf, mg)
    assert mismatches == 0


@pytest.mark.parametrize("mode", [0, 1])
def test_tail_scan_matches(ctx, O, mode):
    fs = 44100
    rng = np.random.default_rng(6 + mode)
    n = fs * 3
    t = np.arange(n) / fs
    x = (0.5 * np.sin(2 * np.pi * 1000 * t) * np.exp(-t * 6.0)).astype(np.float32)
    x = np.stack([x, x * 0.7]).astype(np.float32)
    x += (rng.standard_normal(x.shape) * 10 ** (-96 / 20)).astype(np.float32)
    # window = 2, 8, 1 and 3 whole hops (per-hop partial sums), 9 hops and a non-multiple (windows summed directly)
    for (win, hop, req, start) in ((4410, 2205, 3, 0), (2048, 256, 4, 30000), (2048, 1000, 2, 777), (4410, 2205, 3, n - 100),
                                   (1024, 1024, 2, 5), (3072, 1024, 3, 100), (2304, 256, 2, 40000)):
        for has_nf, nf, mg in ((True, -93.0, -3.0), (True, -96.0, 10.0), (False, 0.0, 10.0)):
            gs, gf = ctx.tail_scan(x, start, win, hop, req, mode, has_nf, nf, mg)
            cs, cf = O.tail_scan(x, start, win, hop, req, mode, has_nf, nf, mg)
            assert gs == cs and np.array_equal(gf, cf), (win, hop, req, start, has_nf, nf, mg)


# ---------------------------------------------------------------- cross-correlation
def test_xcorr_impulse_equals_find_peak(ctx, O):
    rng = np.random.default_rng(8)
    stim = np.array([0.9], np.float32)
    for d in (0, 1, 255, 256, 4097, 65535):
        y = (rng.standard_normal((2, 70000)) * 1e-4).astype(np.float32)
        y[:, d] += 0.8
        g = ctx.xcorr_peak(y, stim, 0, 65536, 0.1)
        c = O.xcorr_peak(y, stim, 0, 65536, 0.1)
        assert g[:3] == c[:3] and g[3] == c[3]
        assert g[1] == ctx.find_peak_position(y[:, :65537], 0.1) == d


def test_xcorr_sweep_and_ties(ctx, O):
    rng = np.random.default_rng(9)
    m = 700
    tt = np.arange(m) / 48000.0
    stim = (0.5 * np.sin(2 * np.pi * (200 + 4000 * tt / tt[-1]) * tt)).astype(np.float32)
    y = (rng.standard_normal((2, 6000)) * 1e-3).astype(np.float32)
    y[0, 1234:1234 + m] += 0.5 * stim
    y[1, 1300:1300 + m] += 0.4 * stim
    g = ctx.xcorr_peak(y, stim, -1024, 4096, 0.01)
    c = O.xcorr_peak(y, stim, -1024, 4096, 0.01)
    assert g == c and g[1] == 1234 and g[2] == 0
    # exact ties: identical copies in both channels and twice in channel 0
    y = np.zeros((2, 6000), np.float32)
    y[1, 100:100 + m] = stim; y[0, 2000:2000 + m] = stim; y[0, 4000:4000 + m] = stim
    g = ctx.xcorr_peak(y, stim, -300, 5000, 0.01)
    c = O.xcorr_peak(y, stim, -300, 5000, 0.01)
    assert g == c and (g[1], g[2]) == (2000, 0)
    # nothing there
    g = ctx.xcorr_peak(np.zeros((2, 512), np.float32), stim, -16, 16, 0.1)
    assert g[0] is False and g[2] == -1
    # negative lag wins
    y = np.zeros((1, 2000), np.float32); y[0, :m - 50] = stim[50:]
    assert ctx.xcorr_peak(y, stim, -200, 200, 0.01) == O.xcorr_peak(y, stim, -200, 200, 0.01)


# ---------------------------------------------------------------- trimLatency
def test_trim_doc_vector(ctx, O):
    cap = rnd((2, 46660), 1)
    g, gn = ctx.trim_latency(cap, 1024, 44100)
    c, cn = O.trim_latency(cap, 1024, 44100)
    assert gn == cn == 44100 and np.array_equal(g, c)


@pytest.mark.parametrize("lat,orig", [(180, 50), (400, 50), (-4, 50), (3, 20), (0, 100), (0, 150), (199, 1), (0, 0)])
def test_trim_edge_cases(ctx, O, lat, orig):
    cap = rnd((2, 100), 2)
    g, gn = ctx.trim_latency(cap, lat, orig)
    c, cn = O.trim_latency(cap, lat, orig)
    assert gn == cn and np.array_equal(g, c)


def test_trim_swift(ctx, O):
    a = rnd(93320, 3)
    for lat, frames, ch in ((1024, 44100, 2), (93000, 44100, 2), (200000, 100, 2), (6, 4, 2), (0, 0, 2)):
        assert np.array_equal(ctx.trim_latency_swift(a, lat, frames, ch), O.trim_latency_swift(a, lat, frames, ch))


def test_remove_dc_reference_order_is_bit_exact(ctx, O):
    """removeDCOffset (Source/MainComponent.cpp:884-902): the reference's ONE float accumulator, reproduced in order => every
    sample bit for bit, short loud buffers and long quiet ones alike."""
    for shape, seed, scale, dc in (((2, 100000), 4, 0.3, 0.01), ((1, 1), 5, 0.3, 0.1), ((3, 4099), 6, 0.5, -0.2), ((2, 7), 7, 1.0, 0.0)):
        x = (rnd(shape, seed, scale) + np.float32(dc)).astype(np.float32)
        assert np.array_equal(ctx.remove_dc_offset(x), O.remove_dc_offset(x)), shape


def test_remove_dc_long_quiet_capture(ctx, O):
    """a 60 s, 44.1 kHz capture with DC 0.003 and 1e-4 of noise (a quiet reverb tail).  The reference's float
    accumulator is ~2.5e-5 off the true mean there -- 26 x 2^-20.  Reference order reproduces the reference exactly; the
    parallel mode subtracts the exactly rounded mean and so differs from the reference by the reference's own drift (recorded)."""
    rng = np.random.default_rng(8)
    x = (0.003 + 1e-4 * rng.standard_normal((2, 60 * 44100))).astype(np.float32)
    ref = O.remove_dc_offset(x)
    assert np.array_equal(ctx.remove_dc_offset(x), ref)                                    # bit exact, drift and all
    par = ctx.remove_dc_offset(x, reference_order=False)
    true_mean = x.astype(np.float64).mean(axis=1, keepdims=True)
    assert np.max(np.abs(par - (x - true_mean.astype(np.float32)))) <= 2.0 ** -23          # the parallel mode removes the true mean
    drift = float(np.max(np.abs((x - ref).astype(np.float64).mean(axis=1) - true_mean[:, 0])))
    gap = float(np.max(np.abs(par - ref)))
    assert abs(gap - drift) <= 2.0 ** -22                                                  # the two modes differ by the reference's drift,
    assert gap > 2.0 ** -20                                                                # which is far above the sample tolerance here
    print(f"reference accumulator drift on the quiet capture: {drift:.3e} ({drift / 2.0 ** -20:.1f} x 2^-20)")


def test_latency_stats_one_pass(ctx, O, f9):
    """findPeakPosition + the sum of squares behind calculateNoiseFloorDb from one read of each capture: the position is the
    reference's (ties, threshold, all-zero buffer), the RMS and the noise floor the reference's floats bit for bit, the peak exact."""
    torch = pytest.importorskip("torch")
    import ctypes as C
    rng = np.random.default_rng(5)
    shapes = [(2, 240000), (1, 50001), (2, 7), (5, 33333), (2, 16384), (2, 16385), (2, 1000), (2, 4096)]
    caps = []
    for i, (ch, frames) in enumerate(shapes):
        x = (rng.standard_normal((ch, frames)) * 10 ** (-80 / 20)).astype(np.float32)
        if i == 6:
            x[:] = 0.0                                                 # all zero: -1
        elif i == 7:
            x[:] *= 100.0                                              # peak below the 0.1 threshold: -1
        else:
            d = int(rng.integers(0, frames))
            x[ch - 1, d] = 0.9
            if i % 2 == 0 and d + 5 < frames:
                x[0, d + 5] = 0.9                                      # equal peak in channel 0, later frame: channel order wins
        caps.append(x)
    d_caps = [torch.from_numpy(c).cuda() for c in caps]
    n = len(caps)
    bufs = (f9.DevBuffer * n)(*[f9.DevBuffer(t.data_ptr(), t.shape[1], t.shape[0], t.shape[1]) for t in d_caps])
    pos = torch.full((n,), -7, dtype=torch.int32, device="cuda")
    pos2 = torch.full((n,), -7, dtype=torch.int32, device="cuda")
    sumsq = torch.zeros(n, dtype=torch.float64, device="cuda")
    peak = torch.zeros(n, dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    ctx._check(f9.lib().f9_dev_latency_stats_batch(ctx.handle, bufs, n, 0.1, pos.data_ptr(), sumsq.data_ptr(), peak.data_ptr()))
    ctx._check(f9.lib().f9_dev_find_peak_batch(ctx.handle, bufs, n, 0.1, pos2.data_ptr()))
    ctx.synchronize()
    pos, pos2, sumsq, peak = pos.cpu().numpy(), pos2.cpu().numpy(), sumsq.cpu().numpy(), peak.cpu().numpy()
    fpp = C.POINTER(C.c_float)
    for x in caps:                                                     # the host-level form: one upload, one read
        rows = [np.ascontiguousarray(r) for r in x]
        chans = (fpp * len(rows))(*[r.ctypes.data_as(fpp) for r in rows])
        p1, nf = C.c_int(-7), C.c_float(0.0)
        ctx._check(f9.lib().f9_measure_latency(ctx.handle, chans, x.shape[0], x.shape[1], 0.1, C.byref(p1), C.byref(nf)))
        assert p1.value == O.find_peak_position(x, 0.1)
        assert np.float32(nf.value).tobytes() == O.noise_floor_db(x).tobytes()
    for i, x in enumerate(caps):
        assert pos[i] == pos2[i] == O.find_peak_position(x, 0.1), i
        assert peak[i] == np.max(np.abs(x))
        rms = np.float32(np.sqrt(sumsq[i] / x.size))
        ref = np.float32(O.calculate_rms(x))
        assert rms.tobytes() == ref.tobytes(), (i, rms, ref)


@pytest.mark.parametrize("remove_dc", [0, 1, 2])
def test_dev_trim_batch_ragged(ctx, O, f9, remove_dc):
    """Device-resident trimLatency (+ fused removeDCOffset) over a ragged batch: latencies at every 16-byte misalignment,
    captures shorter than the request (zero padding), latency past the end, mono / stereo / 5 channels, odd strides.  Trim is
    a bit-exact copy; with DC removal the mean is the parallel double sum (tolerance parity, as test_remove_dc_tolerance)."""
    torch = pytest.importorskip("torch")
    import ctypes as C
    cases = [(2, 50000, 2 * 1001, 40000), (2, 50000, 2 * 1002, 48999), (2, 50000, 2 * 1003, 60000), (1, 9000, 7, 9000),
             (5, 3000, 5 * 13, 2500), (2, 4000, 2 * 5000, 1000), (2, 70001, 2 * 64, 70001), (2, 100, 0, 1)]
    caps = [(rnd((ch, frames), 300 + i, 0.4) + np.float32(0.002 * (i + 1))).astype(np.float32) for i, (ch, frames, _, _) in enumerate(cases)]
    d_caps = [torch.from_numpy(c).cuda() for c in caps]
    d_outs = [torch.full((ch, orig + 3), float("nan"), dtype=torch.float32, device="cuda") for (ch, _, _, orig) in cases]
    n = len(cases)
    cb = (f9.DevBuffer * n)(*[f9.DevBuffer(t.data_ptr(), t.shape[1], t.shape[0], t.shape[1]) for t in d_caps])
    ob = (f9.DevBuffer * n)(*[f9.DevBuffer(t.data_ptr(), t.shape[1], t.shape[0], orig) for t, (_, _, _, orig) in zip(d_outs, cases)])
    lat = (C.c_int * n)(*[c[2] for c in cases])
    torch.cuda.synchronize()
    ctx._check(f9.lib().f9_dev_trim_batch(ctx.handle, cb, lat, ob, n, remove_dc))
    ctx.synchronize()
    for cap, t, (ch, frames, latency, orig) in zip(caps, d_outs, cases):
        got = t.cpu().numpy()
        want, _ = O.trim_latency(cap, latency, orig)
        assert np.all(np.isnan(got[:, orig:]))                       # nothing written past the requested length
        if remove_dc == 2:
            assert np.array_equal(got[:, :orig], O.remove_dc_offset(want))                                  # the reference's accumulator, in order
        elif remove_dc:
            exact = want - (want.astype(np.float64).sum(axis=1, keepdims=True) / orig).astype(np.float32)      # the mean without the float accumulator's drift
            assert np.max(np.abs(got[:, :orig] - exact)) <= 2.0 ** -23
            assert np.max(np.abs(got[:, :orig] - O.remove_dc_offset(want))) <= 2.0 ** -20                     # the reference's sequential float sum
        else:
            assert np.array_equal(got[:, :orig], want)


# ---------------------------------------------------------------- format conversion (bit exact)
@pytest.mark.parametrize("fmt,dt", [(1, np.uint8), (2, np.int16), (3, None), (4, np.int32), (5, np.float32)])
@pytest.mark.parametrize("src_ch,dst_ch,frames", [(1, 2, 1000), (2, 2, 4099), (6, 6, 333), (2, 1, 64)])
def test_pcm_to_planar(ctx, O, fmt, dt, src_ch, dst_ch, frames):
    rng = np.random.default_rng(fmt * 100 + frames)
    nbytes = {1: 1, 2: 2, 3: 3, 4: 4, 5: 4}[fmt] * src_ch * frames
    raw = rng.integers(0, 256, nbytes, dtype=np.uint8)
    if fmt == 5:
        raw = rng.uniform(-1, 1, src_ch * frames).astype(np.float32).view(np.uint8)
    assert np.array_equal(ctx.pcm_to_planar(raw, fmt, src_ch, dst_ch), O.pcm_to_planar(raw, fmt, src_ch, dst_ch))


@pytest.mark.parametrize("ch,frames", [(1, 5), (2, 4099), (2, 44100), (7, 1001), (64, 300)])
def test_planar_to_pcm24(ctx, O, ch, frames):
    x = rnd((ch, frames), ch + frames, 1.2)               # includes clipping on both sides
    x[0, :4] = [1.0, -1.0, 0.0, -0.0][:min(4, frames)] if frames >= 4 else x[0, :4]
    assert np.array_equal(ctx.planar_to_pcm24(x), O.planar_to_pcm24(x))


def test_pcm_batch_forms(ctx, O, f9):
    """One launch for a ragged batch of files (different lengths and channel counts, an empty file, odd byte offsets):
    both directions bit-identical to the oracle's per-file conversion."""
    torch = pytest.importorskip("torch")
    import ctypes as C
    shapes = [(2, 4099), (1, 7), (2, 0), (6, 1001), (2, 44100), (3, 1), (2, 1023), (2, 1024), (2, 1025)]
    xs = [rnd(s, 900 + i, 1.2) for i, s in enumerate(shapes)]
    n = len(xs)
    # planar float -> 24-bit payload
    d_x = [torch.from_numpy(np.ascontiguousarray(x)).cuda() for x in xs]
    d_p = [torch.full((max(1, x.size * 3) + 1,), 0xAB, dtype=torch.uint8, device="cuda") for x in xs]
    bufs = (f9.DevBuffer * n)(*[f9.DevBuffer(t.data_ptr(), x.shape[1], x.shape[0], x.shape[1]) for t, x in zip(d_x, xs)])
    ptrs = (C.c_void_p * n)(*[t.data_ptr() for t in d_p])
    torch.cuda.synchronize()
    ctx._check(f9.lib().f9_dev_planar_to_pcm24_batch(ctx.handle, bufs, ptrs, n))
    ctx.synchronize()
    for x, t in zip(xs, d_p):
        got = t.cpu().numpy()
        assert np.array_equal(got[:x.size * 3], O.planar_to_pcm24(x)) if x.size else True
        assert np.all(got[x.size * 3:] == 0xAB)                       # nothing written past the payload
    # 24-bit payload -> planar float (mono -> stereo duplication through the destination's channel count)
    rng = np.random.default_rng(77)
    frames = [4099, 7, 0, 1001, 44100, 1]
    src_ch = 2
    raws = [rng.integers(0, 256, 3 * src_ch * f, dtype=np.uint8) for f in frames]
    d_raw = [torch.from_numpy(np.concatenate([r, np.zeros(1, np.uint8)])).cuda() for r in raws]
    dst_ch = [2, 1, 2, 3, 2, 2]
    d_out = [torch.full((c, max(f, 1) + 5), float("nan"), dtype=torch.float32, device="cuda") for c, f in zip(dst_ch, frames)]
    bufs = (f9.DevBuffer * len(frames))(*[f9.DevBuffer(t.data_ptr(), t.shape[1], c, f) for t, c, f in zip(d_out, dst_ch, frames)])
    ptrs = (C.c_void_p * len(frames))(*[t.data_ptr() for t in d_raw])
    torch.cuda.synchronize()
    ctx._check(f9.lib().f9_dev_pcm_to_planar_batch(ctx.handle, ptrs, 3, src_ch, bufs, len(frames)))
    ctx.synchronize()
    for r, t, c, f in zip(raws, d_out, dst_ch, frames):
        got = t.cpu().numpy()
        if f:
            assert np.array_equal(got[:, :f], O.pcm_to_planar(r, 3, src_ch, c))
        assert np.all(np.isnan(got[:, f:]))


@pytest.mark.parametrize("ch", [1, 2])
def test_pcm_fast_paths(ctx, O, f9, ch):
    """Uniform mono / stereo batches on 16-byte aligned planes take the 128-bit kernels (a thread owns 16 interleaved samples,
    the payload crosses a warp-private slice of shared memory): ragged lengths around the group size, both directions, 24- and
    16-bit input, mono duplicated into stereo; bit-identical to the oracle and to the byte-staged kernels (F9_PCM_BYTEWISE)."""
    torch = pytest.importorskip("torch")
    import ctypes as C
    frames = [0, 1, 7, 8, 9, 15, 16, 17, 255, 256, 257, 4099, 44100, 441000, 8 * 32 * 8 + 3]
    n = len(frames)
    xs = [rnd((ch, f), 700 + i + ch, 1.2) for i, f in enumerate(frames)]
    for x in xs:
        if x.shape[1] >= 4:
            x[0, :4] = [1.0, -1.0, 0.0, -0.0]
    def pack(c):
        stride = [(f + 3) // 4 * 4 + 4 for f in frames]
        d_x = [torch.zeros((ch, st), dtype=torch.float32, device="cuda") for st in stride]
        for t, x in zip(d_x, xs):
            t[:, :x.shape[1]] = torch.from_numpy(x).cuda()
        d_p = [torch.full((x.size * 3 + 32,), 0xAB, dtype=torch.uint8, device="cuda") for x in xs]
        bufs = (f9.DevBuffer * n)(*[f9.DevBuffer(t.data_ptr(), st, ch, f) for t, st, f in zip(d_x, stride, frames)])
        ptrs = (C.c_void_p * n)(*[t.data_ptr() for t in d_p])
        torch.cuda.synchronize()
        c._check(f9.lib().f9_dev_planar_to_pcm24_batch(c.handle, bufs, ptrs, n))
        c.synchronize()
        return [t.cpu().numpy() for t in d_p]
    got = pack(ctx)
    for x, g in zip(xs, got):
        assert np.array_equal(g[:x.size * 3], O.planar_to_pcm24(x))
        assert np.all(g[x.size * 3:] == 0xAB)
    c2 = f9.Context(0)
    try:
        c2.set_option("F9_PCM_BYTEWISE", 1)
        for a, b in zip(got, pack(c2)):
            assert np.array_equal(a, b)
    finally:
        c2.close()
    # payload -> planes: 24-bit and 16-bit, same channel count and (mono) duplicated into stereo
    rng = np.random.default_rng(78 + ch)
    for fmt, bps in ((3, 3), (2, 2)):
        raws = [rng.integers(0, 256, bps * ch * f, dtype=np.uint8) for f in frames]
        for dst_ch in ([1, 2] if ch == 1 else [2]):
            d_raw = [torch.from_numpy(np.concatenate([r, np.zeros(16, np.uint8)])).cuda() for r in raws]
            stride = [(f + 3) // 4 * 4 + 8 for f in frames]
            d_out = [torch.full((dst_ch, st), float("nan"), dtype=torch.float32, device="cuda") for st in stride]
            bufs = (f9.DevBuffer * n)(*[f9.DevBuffer(t.data_ptr(), st, dst_ch, f) for t, st, f in zip(d_out, stride, frames)])
            ptrs = (C.c_void_p * n)(*[t.data_ptr() for t in d_raw])
            torch.cuda.synchronize()
            ctx._check(f9.lib().f9_dev_pcm_to_planar_batch(ctx.handle, ptrs, fmt, ch, bufs, n))
            ctx.synchronize()
            for r, t, f in zip(raws, d_out, frames):
                g = t.cpu().numpy()
                if f:
                    assert np.array_equal(g[:, :f], O.pcm_to_planar(r, fmt, ch, dst_ch)), (fmt, dst_ch, f)
                assert np.all(np.isnan(g[:, f:]))


@pytest.mark.parametrize("ch,frames", [(1, 5), (2, 4099), (3, 70000), (64, 300)])
def test_interleave_round_trip(ctx, O, ch, frames):
    x = rnd((ch, frames), 10 + ch)
    inter = ctx.interleave(x)
    assert np.array_equal(inter, O.interleave(x))
    assert np.array_equal(ctx.deinterleave(inter, ch), x)
