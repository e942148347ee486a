"""CPU: the oracle against the reference's known-answer vectors and its own invariants.

The vectors in tests/golden/doc_vectors.json are the only pinned numbers the reference holds for this
path (the reference has no test suite; these are worked examples in its docs).
"""
import json
import os

import numpy as np
import pytest

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "doc_vectors.json")))


# ---------------------------------------------------------------- doc vectors
def test_recording_length_vectors(O):
    for v in GOLD["recording_length"]:
        assert O.recording_length(v["src"], v["lat"]) == v["expect"]


def test_latency_frames_and_peak_to_samples(O):
    v = GOLD["latency_frames"]
    assert v["measured"] // v["channels"] == v["expect"]
    # peak at frame 512 of a stereo capture -> measuredLatencySamples 1024 (MainComponent.cpp:275)
    g = GOLD["measured_latency_samples_from_peak_frame"]
    buf = np.zeros((g["channels"], 4096), np.float32)
    buf[:, g["peak_frame"]] = 0.9
    assert O.find_peak_position(buf, 0.1) * 2 == g["expect"]


def test_trim_vector_cpp_and_swift(O):
    t = GOLD["trim"]
    frames = t["captured_samples"] // t["channels"]
    cap = np.arange(t["channels"] * frames, dtype=np.float32).reshape(t["channels"], frames)
    out, copied = O.trim_latency(cap, t["latency_samples"], t["source_frames"])
    assert out.shape == (t["channels"], t["expect_frames"]) and copied == t["expect_frames"]
    assert np.array_equal(out, cap[:, 512:512 + 44100])
    inter = O.interleave(cap)
    sw = O.trim_latency_swift(inter, t["latency_samples"], t["source_frames"], t["channels"])
    assert sw.size == t["expect_samples"]
    assert np.array_equal(sw, inter[t["start_sample"]:t["end_sample"]])
    # both layouts describe the same audio
    assert np.array_equal(O.deinterleave(sw, t["channels"]), out)


def test_threshold_vectors(O):
    for v in GOLD["noise_floor_threshold_db"]:
        assert abs(float(O.noise_floor_threshold_db(True, v["nf"], v["margin"])) - v["expect"]) < 1e-4
    assert float(O.noise_floor_threshold_db(False, -96.0, 10.0)) == GOLD["fallback_threshold_db"]["expect"]


def test_latency_ms_vector(O):
    v = GOLD["latency_ms"]
    assert round(O.latency_ms(v["samples"], v["fs"]), v["places"]) == v["expect"]
    assert O.latency_ms(-1, 44100.0) == 0.0


def test_reverb_window_vector():
    v = GOLD["reverb_window"]
    assert int(v["fs"] * 0.1) == v["window_frames"]


# ---------------------------------------------------------------- findPeakPosition semantics
def test_peak_tie_breaking_and_sentinels(O):
    buf = np.zeros((2, 100), np.float32)
    assert O.find_peak_position(buf, 0.1) == -1                     # all zero
    buf[1, 10] = 0.5; buf[0, 40] = 0.5                               # equal peaks: lower channel wins
    assert O.find_peak_position(buf, 0.1) == 40
    buf[0, 70] = 0.5                                                 # equal peak later in ch0: earliest wins
    assert O.find_peak_position(buf, 0.1) == 40
    buf[1, 5] = -0.6                                                 # strictly larger in ch1 wins, sign ignored
    assert O.find_peak_position(buf, 0.1) == 5
    assert O.find_peak_position(buf, 0.6) == -1                      # max > threshold is strict
    buf[0, 3] = np.nan                                               # NaN never wins
    assert O.find_peak_position(buf, 0.1) == 5


def test_peak_interleaved_swift(O):
    a = np.zeros(1000, np.float32)
    assert O.find_peak_interleaved(a, 0.1) == (0, False)             # default index 0, throws -> not found
    a[700] = 0.3; a[200] = -0.3
    assert O.find_peak_interleaved(a, 0.1) == (200, True)
    assert O.find_peak_interleaved(a, 0.3) == (200, False)


# ---------------------------------------------------------------- RMS / noise floor / tail predicates
def test_rms_and_noise_floor(O):
    rng = np.random.default_rng(1)
    x = rng.uniform(-0.5, 0.5, (2, 2048)).astype(np.float32)
    ref = np.sqrt(np.sum((x * x).astype(np.float64)) / x.size)
    assert abs(float(O.calculate_rms(x)) - ref) <= 2e-7 * ref
    assert float(O.calculate_rms(np.zeros((2, 0), np.float32))) == 0.0
    assert float(O.noise_floor_db(np.zeros((2, 64), np.float32))) == pytest.approx(-120.0, abs=1e-4)   # floor 1e-6


def test_tail_predicates(O):
    w = np.full((2, 2048), 1e-5, np.float32)                         # -100 dB
    assert O.tail_below_floor(w, False, 0.0, 10.0)                   # < -80 fallback
    assert not O.tail_below_floor(w, True, -96.0, 10.0)              # -100 is not < -105.6
    assert O.tail_below_floor(w * 0.5, True, -96.0, 10.0)            # -106.02 < -105.6
    iw = np.full(8820, 5e-5, np.float32)
    assert O.tail_below_floor_swift(iw, False, 0.0, 10.0)            # peak < 1e-4
    assert not O.tail_below_floor_swift(iw * 4, False, 0.0, 10.0)
    assert O.tail_below_floor_swift(np.zeros(8820, np.float32), True, -96.0, 10.0)   # -160 branch


def test_tail_scan_three_consecutive(O):
    fs, hop, win = 44100, 2205, 4410
    n = fs * 2
    x = np.zeros((2, n), np.float32)
    x[:, : fs] = 0.25                                                # loud for 1 s, then silence
    x[:, fs + 3 * hop: fs + 3 * hop + 10] = 0.25                     # a late blip resets the counter
    stop, flags = O.tail_scan(x, 0, win, hop, 3, 1, False, 0.0, 10.0)
    # reproduce the loop by hand
    cons, expect = 0, -1
    for i in range((n - 0) // hop):
        e = (i + 1) * hop
        if e < win:
            assert flags[i] == -1
            continue
        f = int(np.max(np.abs(x[:, e - win:e])) < np.float32(0.0001))
        assert flags[i] == f
        cons = cons + 1 if f else 0
        if cons >= 3 and expect < 0:
            expect = e
    assert stop == expect and stop > fs + 3 * hop


# ---------------------------------------------------------------- trim edge cases
def test_trim_edge_cases(O):
    cap = np.arange(2 * 100, dtype=np.float32).reshape(2, 100) + 1
    out, n = O.trim_latency(cap, 2 * 90, 50)                         # insufficient capture: 10 copied, rest zero
    assert n == 10 and np.array_equal(out[:, :10], cap[:, 90:]) and not out[:, 10:].any()
    out, n = O.trim_latency(cap, 2 * 200, 50)                        # latency past the end
    assert n == 0 and not out.any()
    out, n = O.trim_latency(cap, -4, 50)                             # negative start: zeros (startFrame >= 0 fails)
    assert not out.any()
    out, n = O.trim_latency(cap, 3, 20)                              # odd interleaved latency truncates: 3/2 = 1
    assert np.array_equal(out, cap[:, 1:21])
    sw = O.trim_latency_swift(np.arange(10, dtype=np.float32), 20, 4, 2)     # start >= count: prefix(want)
    assert np.array_equal(sw, np.arange(8, dtype=np.float32))
    sw = O.trim_latency_swift(np.arange(10, dtype=np.float32), 6, 4, 2)      # short: no padding
    assert np.array_equal(sw, np.arange(6, 10, dtype=np.float32))


# ---------------------------------------------------------------- interpolators [JUCE-recall]
RATIOS = [147 / 160, 160 / 147, 320 / 147, 0.25, 2.0, 1.0, 0.731234567]


@pytest.mark.parametrize("kind,lat", [(0, 100), (1, 2), (2, 2), (3, 1), (4, 0)])
def test_interp_latency_and_identity(O, kind, lat):
    it = O.Interpolator(kind)
    assert it.base_latency == lat
    rng = np.random.default_rng(kind)
    x = rng.uniform(-1, 1, 1000).astype(np.float32)
    y, used = it.process(1.0, x, 600)
    assert used == 600
    # ratio 1 from reset: a pure delay of base_latency samples (offset 0 every sample)
    exp = np.concatenate([np.zeros(lat, np.float32), x])[:600]
    tol = 0 if kind in (1, 2, 3, 4) else 1e-6
    assert np.max(np.abs(y - exp)) <= tol


@pytest.mark.parametrize("kind", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("ratio", RATIOS)
def test_interp_streaming_equals_one_shot(O, kind, ratio):
    rng = np.random.default_rng(7)
    n_out = 3000
    x = rng.uniform(-1, 1, int(n_out * ratio) + 300).astype(np.float32)
    a = O.Interpolator(kind)
    whole, used = a.process(ratio, x, n_out)
    b = O.Interpolator(kind)
    parts, pos = [], 0
    for chunk in (1, 7, 256, 1000, n_out - 1264):
        y, u = b.process(ratio, x[pos:], chunk)
        parts.append(y); pos += u
    assert pos == used and a.pos == b.pos
    assert np.array_equal(np.concatenate(parts), whole)


def test_interp_first_output_consumes_one_input(O):
    for kind in range(5):
        it = O.Interpolator(kind)
        _, used = it.process(0.5, np.ones(8, np.float32), 1)
        assert used == 1 and it.pos == 0.5


def test_sinc_sine_snr_against_ideal(O):
    fs_in, fs_out, f0 = 44100, 48000, 1000.0
    n = np.arange(44100)
    x = (0.5 * np.sin(2 * np.pi * f0 * n / fs_in)).astype(np.float32)
    nout = 47000
    y, _ = O.resample_channel(0, fs_in / fs_out, x, nout)
    t = (np.arange(nout) * (fs_in / fs_out) - 100) / fs_in
    ideal = 0.5 * np.sin(2 * np.pi * f0 * t)
    err = (y - ideal)[400:]
    snr = 20 * np.log10(np.sqrt(np.mean(ideal[400:] ** 2)) / np.sqrt(np.mean(err ** 2)))
    assert snr > 120.0


def test_wrap_and_exceeded(O):
    x = np.arange(1, 11, dtype=np.float32)
    it = O.Interpolator(4)                                           # zero-order hold makes the walk visible
    y, used = it.process_wrap(1.0, x, 15, 10, 0)                     # exceeds: zeros after 10 inputs
    assert used == 10 and np.array_equal(y, np.concatenate([x, np.zeros(5, np.float32)]))
    it = O.Interpolator(4)
    y, used = it.process_wrap(1.0, x, 15, 10, 10)                    # wraps around the 10-sample loop
    assert np.array_equal(y, np.concatenate([x, x[:5]])) and used == 5


def test_sinc_table_shape(O):
    t = O.sinc_table()
    assert t.shape == (10001,) and t[0] == 1.0 and t[10000] == 0.0
    assert all(t[100 * k] == 0.0 for k in range(1, 101))
    assert abs(t[1] - 0.999836) < 1e-6                               # T[1] ~ 0.999836


# ---------------------------------------------------------------- cross-correlation definition
def test_xcorr_impulse_equals_find_peak(O):
    rng = np.random.default_rng(3)
    for trial in range(8):
        y = (rng.standard_normal((2, 4000)) * 1e-3).astype(np.float32)
        d = int(rng.integers(0, 1500))
        y[:, d] += 0.7
        stim = np.array([0.9], np.float32)                           # generateImpulse amplitude (MainComponent.cpp:938)
        found, lag, ch, val = O.xcorr_peak(y, stim, 0, 2000, 0.1)
        assert found and lag == O.find_peak_position(y[:, :2001], 0.1) == d


def test_xcorr_ties_and_bounds(O):
    y = np.zeros((2, 300), np.float32)
    stim = np.array([1.0, 0.5, 0.25], np.float32)
    y[1, 50:53] = stim; y[0, 120:123] = stim; y[0, 200:203] = stim   # three equal matches
    found, lag, ch, _ = O.xcorr_peak(y, stim, -64, 250, 0.1)
    assert found and (lag, ch) == (120, 0)                           # lowest channel, then earliest lag
    found, lag, ch, _ = O.xcorr_peak(y, stim, 130, 250, 0.1)
    assert (lag, ch) == (200, 0)
    found, lag, ch, _ = O.xcorr_peak(np.zeros((2, 64), np.float32), stim, -8, 8, 0.1)
    assert not found and ch == -1
    # negative lags: the stimulus starts before the recording
    rstim = stim[::-1].copy()
    y = np.zeros((1, 64), np.float32); y[0, 0:2] = rstim[1:]
    found, lag, ch, _ = O.xcorr_peak(y, rstim, -8, 8, 0.01)
    assert found and lag == -1


# ---------------------------------------------------------------- format conversion [JUCE-recall]
def test_pcm_round_trip_24(O):
    rng = np.random.default_rng(5)
    ints = rng.integers(-(1 << 23), 1 << 23, size=(1000, 2), dtype=np.int64)
    ints[0] = [-(1 << 23), (1 << 23) - 1]
    raw = np.zeros((1000, 2, 3), np.uint8)
    for b in range(3):
        raw[..., b] = (ints >> (8 * b)) & 0xFF
    planar = O.pcm_to_planar(raw.ravel(), O.FMT_S24, 2)
    assert np.array_equal(planar, (ints.T * 256).astype(np.float32) * np.float32(2.0 ** -31))
    back = O.planar_to_pcm24(planar).reshape(1000, 2, 3).astype(np.int64)
    got = back[..., 0] | (back[..., 1] << 8) | (back[..., 2] << 16)
    got = np.where(got >= 1 << 23, got - (1 << 24), got)
    # independent statement of the writer: round-half-even of INT_MAX * (double) x, top 24 bits.  Reading scales by
    # 2^-31 but writing by 2^31 - 1, so the round trip is the identity only for v <= 2^22 (it loses one LSB above).
    exp = (np.rint(2147483647.0 * planar.T.astype(np.float64)).astype(np.int64)) >> 8
    assert np.array_equal(got, exp)
    small = ints <= (1 << 22)
    assert np.array_equal(got[small], ints[small]) and np.all(got[~small] == ints[~small] - 1)


def test_pcm24_clipping_and_rounding(O):
    x = np.array([[1.0, -1.0, 2.0, -3.0, 0.0, 0.5, -0.5, 1e-9]], np.float32)
    b = O.planar_to_pcm24(x).reshape(-1, 3).astype(np.int64)
    v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
    v = np.where(v >= 1 << 23, v - (1 << 24), v)
    assert list(v[:5]) == [8388607, -8388608, 8388607, -8388608, 0]
    assert v[5] == (round(2147483647 * 0.5) >> 8) and v[6] == ((-round(2147483647 * 0.5)) >> 8)


def test_pcm_formats_and_mono_dup(O):
    s16 = np.array([0, 1, -1, 32767, -32768], np.int16)
    p = O.pcm_to_planar(s16.view(np.uint8), O.FMT_S16, 1, 2)          # mono -> stereo duplication
    assert p.shape == (2, 5) and np.array_equal(p[0], p[1])
    assert np.array_equal(p[0], (s16.astype(np.int64) << 16).astype(np.float32) * np.float32(2.0 ** -31))
    u8 = np.array([128, 0, 255], np.uint8)
    assert np.array_equal(O.pcm_to_planar(u8, O.FMT_U8, 1)[0], np.array([0.0, -1.0, 127 / 128], np.float32))
    f32 = np.array([0.25, -0.75], np.float32)
    assert np.array_equal(O.pcm_to_planar(f32.view(np.uint8), O.FMT_F32, 1)[0], f32)


def test_dc_removal(O):
    rng = np.random.default_rng(9)
    x = (rng.standard_normal((2, 5000)) * 0.1 + 0.05).astype(np.float32)
    y = O.remove_dc_offset(x)
    assert np.all(np.abs(y.mean(axis=1)) < 1e-6)


def test_stimuli(O):
    imp = O.generate_impulse(2, 16)
    assert imp[0, 0] == np.float32(0.9) and imp[1, 0] == np.float32(0.9) and not imp[:, 1:].any()
    s, ph = O.generate_sine(2, 441, 1000.0, 44100.0, 0.0)
    assert np.max(np.abs(s)) <= 0.5 and np.array_equal(s[0], s[1]) and s[0, 0] == 0.0
