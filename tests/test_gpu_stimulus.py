"""GPU parity of the stimuli: generateImpulse (Source/MainComponent.cpp:934-945) and generateSineWave (:907-932, callback
variant :141-167, Swift SineWaveGenerator.swift:35-59) through the C ABI against the oracle.

The impulse and the phase chain (float in C++, double in Swift) are bit-exact.  The samples are amplitude * sin(phase): the
device rounds the double-precision sine to float (the correctly rounded value); std::sin(float) is the platform libm's sinf
(glibc here, Apple's libm where the reference runs), which is within one float ulp of it and differs on 1-2 % of arguments."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def ulp_diff(a, b):
    ai = a.view(np.int32).astype(np.int64); bi = b.view(np.int32).astype(np.int64)
    ai = np.where(ai < 0, -(ai & 0x7fffffff), ai); bi = np.where(bi < 0, -(bi & 0x7fffffff), bi)
    return np.abs(ai - bi)


def _phases(n, fs, block, O):
    """float phase chain of block `block` (numpy float32 restatement of Source/MainComponent.cpp:916-926, :929-931)"""
    f = np.float32
    two_pi = f(2.0) * f(np.pi)
    inc = f(f(f(1000.0) * f(2.0)) * f(np.pi)) / f(fs)
    start = f(0.0)
    for _ in range(block):
        start = f(start + f(inc * f(n)))
        if start >= two_pi:
            start = f(start - two_pi)
    out = np.empty(n, np.float32)
    ph = start
    for i in range(n):
        out[i] = ph
        ph = f(ph + inc)
        if ph >= two_pi:
            ph = f(ph - two_pi)
    return out


@pytest.mark.parametrize("shape", [(1, 1), (2, 128), (2, 1024), (3, 48000), (8, 7)])
def test_generate_impulse(ctx, O, shape):
    got = ctx.generate_impulse(*shape)
    assert np.array_equal(got, O.generate_impulse(*shape))
    assert got[0, 0] == np.float32(0.9)
    # the latency measurement's own loop: the impulse peak-picks at frame 0
    assert ctx.find_peak_position(got, 0.1) == O.find_peak_position(got, 0.1) == 0


@pytest.mark.parametrize("n,fs", [(128, 44100.0), (1024, 48000.0), (441, 44100.0), (96000, 96000.0), (220500, 44100.0)])
def test_generate_sine_wave_blocks(ctx, O, n, fs):
    ph_g = ph_o = 0.0
    for block in range(3):                                   # sinePhase carried from block to block (closed-form update, :929-931)
        got, ph_g = ctx.generate_sine_wave(2, n, ph_g, 1000.0, fs)
        ref, ph_o = O.generate_sine(2, n, 1000.0, fs, ph_o)
        assert ph_g.tobytes() == ph_o.tobytes()              # the member's update, bit for bit
        d = ulp_diff(got, ref)
        assert d.max() <= 1, f"block {block}: {d.max()} ulp"
        assert (d != 0).mean() < 0.05                        # glibc's sinf misses the correctly rounded value on 1-2 % of arguments
        exact = (0.5 * np.sin(_phases(n, fs, block, O).astype(np.float64)).astype(np.float32)).astype(np.float32)
        assert np.array_equal(got[0], exact)                 # the device's samples ARE the correctly rounded ones
        assert np.array_equal(got[0], got[1])
        assert np.abs(got).max() <= 0.5


def test_generate_sine_wave_callback_form(ctx, O):
    ph_g = ph_o = 0.0
    for n in (512, 512, 333, 48000):                         # device blocks of the audio callback (:141-167): the chain is the member
        got, ph_g = ctx.generate_sine_wave(2, n, ph_g, 1000.0, 48000.0, callback_form=True)
        ref, ph_o = O.generate_sine_callback(2, n, 1000.0, 48000.0, ph_o)
        assert ph_g.tobytes() == ph_o.tobytes()
        assert ulp_diff(got, ref).max() <= 1


def test_generate_sine_wave_swift(ctx, O):
    ph_g = ph_o = 0.0
    for frames, ch in ((512, 2), (4410, 2), (1000, 1), (30000, 4)):
        got, ph_g = ctx.generate_sine_wave_swift(frames, ch, ph_g, 1000.0, 44100.0, 0.5)
        ref, ph_o = O.generate_sine_swift(frames, ch, 1000.0, 44100.0, 0.5, ph_o)
        assert ph_g == ph_o                                  # double chain, bit for bit
        assert ulp_diff(got, ref).max() <= 1
        assert np.array_equal(got.reshape(frames, ch)[:, 0], got.reshape(frames, ch)[:, -1])


def test_generate_edge_cases(ctx, f9):
    assert ctx.generate_impulse(2, 0).shape == (2, 0)
    out, ph = ctx.generate_sine_wave(2, 0, 1.25)
    assert out.shape == (2, 0) and ph == np.float32(1.25)
    with pytest.raises(f9.F9Error):
        ctx.generate_sine_wave(2, 16, 0.0, 1000.0, 0.0)
