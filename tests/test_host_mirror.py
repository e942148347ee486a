"""The C++ host mirror (host/F9Dsp.hpp): it must compile against the C ABI with plain g++ (CPU test) and, on a GPU,
give the oracle's answers through the reference-shaped calls (gpu test)."""
import os
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "f9-juce-resampler-studio_b200")
EXE = os.path.join(PKG, "build", "host_mirror_main")


def build_exe():
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(PKG, "host"),
                           os.path.join(ROOT, "tests", "cpp", "host_mirror_main.cpp"), "-o", EXE,
                           "-L", os.path.join(PKG, "lib"), "-lf9dsp", "-Wl,-rpath," + os.path.join(PKG, "lib")])


def test_host_mirror_compiles_with_plain_gxx(f9):
    f9.lib()
    build_exe()
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_host_mirror_matches_oracle(O, tmp_path):
    build_exe()
    rng = np.random.default_rng(0)
    frames, playback, lat = 60000, 44100, 777
    cap = (rng.standard_normal((2, frames)) * 1e-4 + 0.003).astype(np.float32)
    cap[:, lat] += 0.9
    inp, outp = tmp_path / "in.bin", tmp_path / "out.bin"
    with open(inp, "wb") as f:
        f.write(struct.pack("<ii", frames, playback)); f.write(cap.tobytes())
    subprocess.check_call([EXE, str(inp), str(outp)])
    raw = open(outp, "rb").read()
    hdr = struct.unpack("<8i", raw[:32]); rms, nf = struct.unpack("<2f", raw[32:40])
    body = np.frombuffer(raw[40:], np.float32)
    trimmed = body[: 2 * playback].reshape(2, playback)
    trimmed_only = body[2 * playback: 4 * playback].reshape(2, playback)
    o1, o2 = body[4 * playback: 4 * playback + 4000], body[4 * playback + 4000: 4 * playback + 8000]
    ras = body[4 * playback + 8000: 4 * playback + 12000].reshape(2, 2000)
    rest = body[4 * playback + 12000:]
    imp, sine, phase = rest[:512].reshape(2, 256), rest[512: 512 + 2048].reshape(2, 1024), rest[512 + 2048]
    assert hdr[0] == 1 and hdr[1] == 2 * O.find_peak_position(cap, 0.1) == 2 * lat
    assert np.float32(nf).tobytes() == O.noise_floor_db(cap).tobytes() and np.float32(rms).tobytes() == O.calculate_rms(cap).tobytes()
    t, _ = O.trim_latency(cap, 2 * lat, playback)
    assert np.array_equal(trimmed_only, t)
    assert np.array_equal(trimmed, O.remove_dc_offset(t))            # removeDCOffset: the reference's own accumulator
    assert hdr[2] == int(O.tail_below_floor(t, True, float(np.float32(nf)), 10.0))
    y1, u1 = O.Interpolator(0).process(0.91875, cap[0], 4000)
    y2, u2 = O.Interpolator(1).process(0.91875, cap[1], 4000)
    assert (hdr[3], hdr[4], hdr[5], hdr[6]) == (u1, u2, 4000, playback)
    assert hdr[7] == O.recording_length(playback, lat)
    assert np.max(np.abs(o1 - y1)) <= 2.0 ** -20 and np.max(np.abs(o2 - y2)) <= 2.0 ** -20
    # juce::ResamplingAudioSource through the AudioSource-shaped mirror: bit-exact against the oracle's object
    src = O.ResamplingAudioSource(cap)
    src.set_resampling_ratio(96000.0 / 44100.0)
    src.prepare_to_play(512)
    ref = np.concatenate([src.get_next_audio_block(n) for n in (512, 333, 1155)], axis=1)
    assert np.array_equal(ras, ref)
    # generateImpulse / generateSineWave through the MainComponent-shaped members
    assert np.array_equal(imp, O.generate_impulse(2, 256))
    a, ph = O.generate_sine(2, 512, 1000.0, 44100.0, 0.0)
    b, ph = O.generate_sine(2, 512, 1000.0, 44100.0, float(ph))
    assert phase.tobytes() == ph.tobytes()
    assert np.max(np.abs(sine - np.concatenate([a, b], axis=1))) <= 2.0 ** -24
