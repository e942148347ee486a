"""CPU: the C-ABI library loads, exports every symbol include/f9dsp.h declares, its host scalars match the
reference's documented vectors, and compute entry points fail loudly when there is no GPU (no CPU fallback)."""
import ctypes as C
import json
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "doc_vectors.json")))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "f9dsp.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"F9_API\s+[^;(]*?\b(f9_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol(f9):
    L = f9.lib()
    syms = header_symbols()
    assert len(syms) >= 50
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/f9dsp.h but not exported by libf9dsp.so"
    bound = {name for name, _, _ in f9.SYMBOLS}
    assert bound == set(syms), (bound ^ set(syms))


def test_library_has_sm100a_code_only(f9):
    out = subprocess.run(["cuobjdump", "--list-elf", f9.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_oracle_in_product():
    """The product may not import, link or call the oracle."""
    pkg = os.path.join(ROOT, "f9-juce-resampler-studio_b200")
    for d, _, files in os.walk(pkg):
        if os.path.basename(d) in ("build", "lib", "__pycache__"):
            continue
        for fn in files:
            if fn.endswith((".cu", ".cuh", ".cpp", ".h", ".hpp", ".py", "Makefile")):
                src = open(os.path.join(d, fn), errors="ignore").read()
                code = re.sub(r"//.*|#.*", "", src) if fn.endswith(".py") is False else re.sub(r"#.*", "", src)
                assert "f9oracle" not in code and "orc_" not in code and "from oracle" not in code and "import oracle" not in code, fn
    ldd = subprocess.run(["ldd", os.path.join(pkg, "lib", "libf9dsp.so")], capture_output=True, text=True).stdout
    assert "oracle" not in ldd


def test_settings_math_doc_vectors(f9):
    for v in GOLD["recording_length"]:
        assert f9.recording_length(v["src"], v["lat"]) == v["expect"]
    for v in GOLD["noise_floor_threshold_db"]:
        assert abs(float(f9.noise_floor_threshold_db(True, v["nf"], v["margin"])) - v["expect"]) < 1e-4
    assert float(f9.noise_floor_threshold_db(False, -96.0, 10.0)) == GOLD["fallback_threshold_db"]["expect"]
    v = GOLD["latency_ms"]
    assert round(f9.latency_ms(v["samples"], v["fs"]), v["places"]) == v["expect"]


def test_settings_math_equals_oracle(f9, O):
    rng = np.random.default_rng(0)
    for _ in range(200):
        nf, mg, db = float(rng.uniform(-130, -20)), float(rng.choice(np.arange(0, 55, 5))), float(rng.uniform(-60, -20))
        assert f9.noise_floor_threshold_db(True, nf, mg) == O.noise_floor_threshold_db(True, nf, mg)
        assert f9.threshold_linear(db) == O.threshold_linear(db)
        s, l = int(rng.integers(0, 10_000_000)), int(rng.integers(0, 100_000))
        assert f9.recording_length(s, l) == O.recording_length(s, l)
        assert f9.latency_ms(l, 44100.0) == O.latency_ms(l, 44100.0)
    assert f9.needs_latency_remeasurement(-1, 256, 256) and not f9.needs_latency_remeasurement(1024, 256, 256)
    assert f9.needs_latency_remeasurement(1024, 256, 512) == O.needs_latency_remeasurement(1024, 256, 512)


def test_default_sinc_table_equals_oracle(f9, O):
    assert np.array_equal(f9.default_sinc_table(), O.sinc_table())


def test_resampled_length(f9):
    assert f9.resampled_length(2_646_000, 44100, 48000) == 2_880_000          # config 1
    assert f9.resampled_length(960_000, 96000, 44100) == 441_000              # config 2
    assert f9.resampled_length(28_800_000, 48000, 192000) == 115_200_000      # config 3
    assert f9.resampled_length(30011, 96000, 44100) == -((-30011 * 147) // 320)
    assert f9.resampled_length(1000, 44100, 44100) == 1000 and f9.resampled_length(0, 44100, 48000) == 0


def test_segment_input_range(f9, O):
    """The halo a time segment needs: 199 inputs before its first fresh one for WindowedSinc, 4 for Lagrange."""
    for kind, taps in ((0, 200), (1, 5)):
        for ratio in (147 / 160, 320 / 147, 0.25, 0.731234567):
            for n0, cnt in ((0, 100), (12345, 1000), (10 ** 8, 4096)):
                first, last = f9.segment_input_range(kind, ratio, n0, cnt)
                exact_first = int(np.floor(float(n0) * ratio + 1e-12))        # newest input of output n0 (approx.)
                assert abs((first + taps - 1) - exact_first) <= 2
                assert last - first >= int(cnt * ratio) + taps - 2 and last - first <= int(cnt * ratio) + taps + 4
    # the oracle reading exactly that range reproduces the segment (checked on the GPU in test_gpu_resample)


def test_context_creation_fails_loudly_without_gpu(f9):
    if f9.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(f9.F9Error) as ei:
        f9.Context(0)
    assert ei.value.code == f9.ERR_NO_DEVICE
    # and the raw ABI reports it too
    h = C.c_void_p(None)
    assert f9.lib().f9_context_create(0, C.byref(h)) == f9.ERR_NO_DEVICE and not h.value
    assert f9.lib().f9_last_error(None)


def test_null_context_is_rejected(f9):
    L = f9.lib()
    out = C.c_int(0)
    assert L.f9_find_peak_position(None, None, 0, 0, 0.1, C.byref(out)) == f9.ERR_INVALID
    assert L.f9_resample_plan_run(None) == f9.ERR_INVALID
    assert L.f9_process_batch(None, None, 0, None) == f9.ERR_INVALID
