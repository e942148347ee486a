"""ctypes front-end of the CPU oracle (oracle/f9_oracle.cpp).

TEST INFRASTRUCTURE ONLY: tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs load this; the product (f9-juce-resampler-studio_b200/) never does.
Every wrapper names the oracle entry point; the reference file:line it follows is on the
C++ function.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libf9oracle.so")

KIND_SINC, KIND_LAGRANGE, KIND_CATMULL, KIND_LINEAR, KIND_ZOH = 0, 1, 2, 3, 4
FMT_U8, FMT_S16, FMT_S24, FMT_S32, FMT_F32 = 1, 2, 3, 4, 5
_BYTES = {FMT_U8: 1, FMT_S16: 2, FMT_S24: 3, FMT_S32: 4, FMT_F32: 4}


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "f9_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _LIB_PATH


_lib = None
_fp = C.POINTER(C.c_float)
_fpp = C.POINTER(_fp)


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_find_peak_position.restype = C.c_int
        L.orc_find_peak_position.argtypes = [_fpp, C.c_int, C.c_int, C.c_float]
        L.orc_find_peak_interleaved.restype = C.c_longlong
        L.orc_find_peak_interleaved.argtypes = [_fp, C.c_longlong, C.c_float, C.POINTER(C.c_int)]
        L.orc_calculate_rms.restype = C.c_float
        L.orc_calculate_rms.argtypes = [_fpp, C.c_int, C.c_int]
        L.orc_noise_floor_db.restype = C.c_float
        L.orc_noise_floor_db.argtypes = [_fpp, C.c_int, C.c_int]
        L.orc_noise_floor_db_swift.restype = C.c_float
        L.orc_noise_floor_db_swift.argtypes = [_fp, C.c_longlong]
        L.orc_recording_length.restype = C.c_int
        L.orc_recording_length.argtypes = [C.c_int, C.c_int]
        L.orc_threshold_linear.restype = C.c_float
        L.orc_threshold_linear.argtypes = [C.c_float]
        L.orc_noise_floor_threshold_db.restype = C.c_float
        L.orc_noise_floor_threshold_db.argtypes = [C.c_int, C.c_float, C.c_float]
        L.orc_latency_ms.restype = C.c_double
        L.orc_latency_ms.argtypes = [C.c_int, C.c_double]
        L.orc_needs_latency_remeasurement.restype = C.c_int
        L.orc_needs_latency_remeasurement.argtypes = [C.c_int, C.c_int, C.c_int]
        L.orc_trim_latency.restype = C.c_int
        L.orc_trim_latency.argtypes = [_fpp, C.c_int, C.c_int, C.c_int, C.c_int, _fpp]
        L.orc_trim_latency_swift.restype = C.c_longlong
        L.orc_trim_latency_swift.argtypes = [_fp, C.c_longlong, C.c_longlong, C.c_longlong, C.c_int, _fp]
        L.orc_tail_below_floor.restype = C.c_int
        L.orc_tail_below_floor.argtypes = [_fpp, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float]
        L.orc_tail_below_floor_swift.restype = C.c_int
        L.orc_tail_below_floor_swift.argtypes = [_fp, C.c_longlong, C.c_int, C.c_float, C.c_float]
        L.orc_tail_scan.restype = C.c_longlong
        L.orc_tail_scan.argtypes = [_fpp, C.c_int, C.c_longlong, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int,
                                    C.c_int, C.c_float, C.c_float, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.orc_remove_dc_offset.restype = None
        L.orc_remove_dc_offset.argtypes = [_fpp, C.c_int, C.c_int]
        L.orc_generate_impulse.restype = None
        L.orc_generate_impulse.argtypes = [_fpp, C.c_int, C.c_int]
        L.orc_generate_sine.restype = C.c_float
        L.orc_generate_sine.argtypes = [_fpp, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float]
        L.orc_generate_sine_callback.restype = C.c_float
        L.orc_generate_sine_callback.argtypes = [_fpp, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float]
        L.orc_generate_sine_swift.restype = C.c_double
        L.orc_generate_sine_swift.argtypes = [_fp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_float, C.c_double]
        L.orc_resample_channel_exact.restype = C.c_int
        L.orc_resample_channel_exact.argtypes = [_fp, C.c_double, _fp, C.c_int, C.POINTER(C.c_double), C.c_int]
        L.orc_sinc_table.restype = None
        L.orc_sinc_table.argtypes = [_fp]
        L.orc_interp_create.restype = C.c_void_p
        L.orc_interp_create.argtypes = [C.c_int, _fp]
        L.orc_interp_destroy.restype = None
        L.orc_interp_destroy.argtypes = [C.c_void_p]
        L.orc_interp_reset.restype = None
        L.orc_interp_reset.argtypes = [C.c_void_p]
        L.orc_interp_latency.restype = C.c_float
        L.orc_interp_latency.argtypes = [C.c_void_p]
        L.orc_interp_pos.restype = C.c_double
        L.orc_interp_pos.argtypes = [C.c_void_p]
        L.orc_interp_process.restype = C.c_int
        L.orc_interp_process.argtypes = [C.c_void_p, C.c_double, _fp, _fp, C.c_int]
        L.orc_interp_process_adding.restype = C.c_int
        L.orc_interp_process_adding.argtypes = [C.c_void_p, C.c_double, _fp, _fp, C.c_int, C.c_float]
        L.orc_interp_process_wrap.restype = C.c_int
        L.orc_interp_process_wrap.argtypes = [C.c_void_p, C.c_double, _fp, _fp, C.c_int, C.c_int, C.c_int]
        L.orc_resample_channel.restype = C.c_int
        L.orc_resample_channel.argtypes = [C.c_int, _fp, C.c_double, _fp, C.c_int, _fp, C.c_int]
        L.orc_resample_channels.restype = None
        L.orc_resample_channels.argtypes = [C.c_int, _fp, C.c_double, _fpp, C.c_int, _fpp, C.c_int, C.c_int, C.c_int]
        L.orc_xcorr_peak.restype = C.c_int
        L.orc_xcorr_peak.argtypes = [_fpp, C.c_int, C.c_int, _fp, C.c_int, C.c_int, C.c_int, C.c_float,
                                     C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_double)]
        L.orc_xcorr_at.restype = C.c_double
        L.orc_xcorr_at.argtypes = [_fp, C.c_int, _fp, C.c_int, C.c_int]
        L.orc_pcm_to_planar.restype = None
        L.orc_pcm_to_planar.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_longlong, _fpp, C.c_int]
        L.orc_planar_to_pcm24.restype = None
        L.orc_planar_to_pcm24.argtypes = [_fpp, C.c_int, C.c_longlong, C.c_void_p]
        L.orc_interleave.restype = None
        L.orc_interleave.argtypes = [_fpp, C.c_int, C.c_longlong, _fp]
        L.orc_deinterleave.restype = None
        L.orc_deinterleave.argtypes = [_fp, C.c_int, C.c_longlong, _fpp]
        L.orc_resampling_source_coeffs.restype = None
        L.orc_resampling_source_coeffs.argtypes = [C.c_double, C.POINTER(C.c_double)]
        L.orc_ras_create.restype = C.c_void_p
        L.orc_ras_create.argtypes = [C.c_int, C.c_int]
        L.orc_ras_destroy.restype = None
        L.orc_ras_destroy.argtypes = [C.c_void_p]
        L.orc_ras_set_ratio.restype = None
        L.orc_ras_set_ratio.argtypes = [C.c_void_p, C.c_double]
        L.orc_ras_prepare.restype = None
        L.orc_ras_prepare.argtypes = [C.c_void_p, C.c_int]
        L.orc_ras_flush.restype = None
        L.orc_ras_flush.argtypes = [C.c_void_p]
        L.orc_ras_get_next_block.restype = C.c_longlong
        L.orc_ras_get_next_block.argtypes = [C.c_void_p, _fpp, C.c_longlong, C.c_longlong, _fpp, C.c_int]
        L.orc_ras_convert.restype = None
        L.orc_ras_convert.argtypes = [C.c_int, _fpp, C.c_longlong, C.c_double, _fpp, C.c_longlong, C.c_int, C.c_int]
        _lib = L
    return _lib


# ------------------------------------------------------------------ helpers
def _planar(a: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float32)
    if a.ndim == 1:
        a = a[None, :]
    assert a.ndim == 2
    return a


def _chan_ptrs(a: np.ndarray):
    """float** over the rows of a C-contiguous 2-D float32 array."""
    n = a.shape[0]
    arr = (_fp * max(n, 1))()
    for c in range(n):
        arr[c] = C.cast(a.ctypes.data + c * a.strides[0], _fp)
    return arr


def _p(a: np.ndarray):
    return a.ctypes.data_as(_fp)


# ------------------------------------------------------------------ detection
def find_peak_position(buf, threshold: float) -> int:
    a = _planar(buf)
    return lib().orc_find_peak_position(_chan_ptrs(a), a.shape[0], a.shape[1], threshold)


def find_peak_interleaved(audio, threshold: float):
    a = np.ascontiguousarray(audio, dtype=np.float32).ravel()
    found = C.c_int(0)
    idx = lib().orc_find_peak_interleaved(_p(a), a.size, threshold, C.byref(found))
    return int(idx), bool(found.value)


def calculate_rms(buf) -> np.float32:
    a = _planar(buf)
    return np.float32(lib().orc_calculate_rms(_chan_ptrs(a), a.shape[0], a.shape[1]))


def noise_floor_db(buf) -> np.float32:
    a = _planar(buf)
    return np.float32(lib().orc_noise_floor_db(_chan_ptrs(a), a.shape[0], a.shape[1]))


def noise_floor_db_swift(audio) -> np.float32:
    a = np.ascontiguousarray(audio, dtype=np.float32).ravel()
    return np.float32(lib().orc_noise_floor_db_swift(_p(a), a.size))


# ------------------------------------------------------------------ settings math
def recording_length(src: int, lat: int) -> int:
    return lib().orc_recording_length(src, lat)


def threshold_linear(db: float) -> np.float32:
    return np.float32(lib().orc_threshold_linear(db))


def noise_floor_threshold_db(has_nf: bool, nf_db: float, margin_pct: float) -> np.float32:
    return np.float32(lib().orc_noise_floor_threshold_db(int(has_nf), nf_db, margin_pct))


def latency_ms(samples: int, fs: float) -> float:
    return lib().orc_latency_ms(samples, fs)


def needs_latency_remeasurement(measured: int, last_buf: int, cur_buf: int) -> bool:
    return bool(lib().orc_needs_latency_remeasurement(measured, last_buf, cur_buf))


# ------------------------------------------------------------------ trim / tail
def trim_latency(captured, latency_samples: int, original_length: int):
    a = _planar(captured)
    out = np.empty((a.shape[0], max(original_length, 0)), dtype=np.float32)
    n = lib().orc_trim_latency(_chan_ptrs(a), a.shape[0], a.shape[1], latency_samples, original_length, _chan_ptrs(out))
    return out, n


def trim_latency_swift(captured, latency_samples: int, source_frames: int, channels: int) -> np.ndarray:
    a = np.ascontiguousarray(captured, dtype=np.float32).ravel()
    out = np.empty(max(source_frames * channels, 1), dtype=np.float32)
    n = lib().orc_trim_latency_swift(_p(a), a.size, latency_samples, source_frames, channels, _p(out))
    return out[:n].copy()


def tail_below_floor(window, has_nf: bool, nf_db: float, margin_pct: float) -> bool:
    a = _planar(window)
    return bool(lib().orc_tail_below_floor(_chan_ptrs(a), a.shape[0], a.shape[1], int(has_nf), nf_db, margin_pct))


def tail_below_floor_swift(window, has_nf: bool, nf_db: float, margin_pct: float) -> bool:
    a = np.ascontiguousarray(window, dtype=np.float32).ravel()
    return bool(lib().orc_tail_below_floor_swift(_p(a), a.size, int(has_nf), nf_db, margin_pct))


def tail_scan(buf, start_frame: int, window: int, hop: int, required: int, mode: int,
              has_nf: bool, nf_db: float, margin_pct: float):
    """Returns (stop_frame or -1, flags[int32 per poll])."""
    a = _planar(buf)
    max_polls = max(0, (a.shape[1] - start_frame) // max(hop, 1)) + 1
    flags = np.full(max_polls, -2, dtype=np.int32)
    npolls = C.c_int(0)
    stop = lib().orc_tail_scan(_chan_ptrs(a), a.shape[0], a.shape[1], start_frame, window, hop, required, mode,
                               int(has_nf), nf_db, margin_pct, flags.ctypes.data_as(C.POINTER(C.c_int)), C.byref(npolls))
    return int(stop), flags[: npolls.value].copy()


def remove_dc_offset(buf) -> np.ndarray:
    a = _planar(buf).copy()
    lib().orc_remove_dc_offset(_chan_ptrs(a), a.shape[0], a.shape[1])
    return a


# ------------------------------------------------------------------ stimuli
def generate_impulse(num_ch: int, num_frames: int) -> np.ndarray:
    a = np.empty((num_ch, num_frames), dtype=np.float32)
    lib().orc_generate_impulse(_chan_ptrs(a), num_ch, num_frames)
    return a


def generate_sine(num_ch: int, n: int, freq: float = 1000.0, fs: float = 44100.0, phase: float = 0.0):
    a = np.empty((num_ch, n), dtype=np.float32)
    ph = lib().orc_generate_sine(_chan_ptrs(a), num_ch, n, freq, fs, phase)
    return a, np.float32(ph)


def generate_sine_callback(num_ch: int, n: int, freq: float = 1000.0, fs: float = 44100.0, phase: float = 0.0):
    a = np.empty((num_ch, n), dtype=np.float32)
    ph = lib().orc_generate_sine_callback(_chan_ptrs(a), num_ch, n, freq, fs, phase)
    return a, np.float32(ph)


def generate_sine_swift(frames: int, channels: int, freq: float = 1000.0, fs: float = 44100.0, amplitude: float = 0.5, phase: float = 0.0):
    a = np.empty(max(frames * channels, 1), dtype=np.float32)
    ph = lib().orc_generate_sine_swift(a.ctypes.data_as(_fp), frames, channels, freq, fs, amplitude, phase)
    return a[: frames * channels], ph


# ------------------------------------------------------------------ interpolators
def sinc_table() -> np.ndarray:
    t = np.empty(10001, dtype=np.float32)
    lib().orc_sinc_table(_p(t))
    return t


class Interpolator:
    """juce::Interpolators-shaped stateful object: process(ratio, in, numOut) -> (out, numUsed)."""

    def __init__(self, kind: int, table: np.ndarray | None = None):
        self._table = None if table is None else np.ascontiguousarray(table, dtype=np.float32)
        self._h = lib().orc_interp_create(kind, None if self._table is None else _p(self._table))
        if not self._h:
            raise ValueError("bad interpolator kind")

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_interp_destroy(self._h)
            self._h = None

    def reset(self):
        lib().orc_interp_reset(self._h)

    @property
    def base_latency(self) -> float:
        return lib().orc_interp_latency(self._h)

    @property
    def pos(self) -> float:
        return lib().orc_interp_pos(self._h)

    def process(self, ratio: float, inp, num_out: int):
        a = np.ascontiguousarray(inp, dtype=np.float32)
        out = np.empty(num_out, dtype=np.float32)
        used = lib().orc_interp_process(self._h, ratio, _p(a), _p(out), num_out)
        assert used <= a.size, "oracle read past the supplied input"
        return out, used

    def process_adding(self, ratio: float, inp, out: np.ndarray, gain: float) -> int:
        a = np.ascontiguousarray(inp, dtype=np.float32)
        assert out.dtype == np.float32 and out.flags.c_contiguous
        return lib().orc_interp_process_adding(self._h, ratio, _p(a), _p(out), out.size, gain)

    def process_wrap(self, ratio: float, inp, num_out: int, avail: int, wrap: int):
        a = np.ascontiguousarray(inp, dtype=np.float32)
        out = np.empty(num_out, dtype=np.float32)
        used = lib().orc_interp_process_wrap(self._h, ratio, _p(a), _p(out), num_out, avail, wrap)
        return out, used


def resample_channel(kind: int, ratio: float, inp, num_out: int, table=None):
    a = np.ascontiguousarray(inp, dtype=np.float32)
    out = np.empty(max(num_out, 0), dtype=np.float32)
    t = None if table is None else np.ascontiguousarray(table, dtype=np.float32)
    used = lib().orc_resample_channel(kind, None if t is None else _p(t), ratio, _p(a), a.size, _p(out), num_out)
    return out, used


def resample_channel_exact(ratio: float, inp, num_out: int, table=None) -> np.ndarray:
    """WindowedSinc with the 200-term sum in double (the interpolator's own float weights and positions): the exact value the
    float paths approximate -- a yardstick, not the reference's output."""
    a = np.ascontiguousarray(inp, dtype=np.float32)
    out = np.empty(max(num_out, 0), dtype=np.float64)
    t = None if table is None else np.ascontiguousarray(table, dtype=np.float32)
    lib().orc_resample_channel_exact(None if t is None else _p(t), ratio, _p(a), a.size, out.ctypes.data_as(C.POINTER(C.c_double)), num_out)
    return out


def resample_channels(kind: int, ratio: float, buf, num_out: int, threads: int = 1, table=None) -> np.ndarray:
    """CPU baseline: one interpolator per channel, channels split statically over host threads."""
    import threading

    a = _planar(buf)
    out = np.empty((a.shape[0], num_out), dtype=np.float32)
    t = None if table is None else np.ascontiguousarray(table, dtype=np.float32)
    ip, op = _chan_ptrs(a), _chan_ptrs(out)
    nch = a.shape[0]
    threads = max(1, min(threads, nch))
    bounds = [nch * i // threads for i in range(threads + 1)]

    def work(c0, c1):
        lib().orc_resample_channels(kind, None if t is None else _p(t), ratio, ip, a.shape[1], op, num_out, c0, c1)

    if threads == 1:
        work(0, nch)
    else:
        ts = [threading.Thread(target=work, args=(bounds[i], bounds[i + 1])) for i in range(threads)]
        [x.start() for x in ts]
        [x.join() for x in ts]
    return out


# ------------------------------------------------------------------ cross-correlation
def xcorr_peak(y, x, lag_min: int, lag_max: int, threshold: float):
    """Returns (found, lag, channel, |r|max)."""
    a = _planar(y)
    s = np.ascontiguousarray(x, dtype=np.float32).ravel()
    lag, ch, val = C.c_int(0), C.c_int(0), C.c_double(0)
    found = lib().orc_xcorr_peak(_chan_ptrs(a), a.shape[0], a.shape[1], _p(s), s.size, lag_min, lag_max, threshold,
                                 C.byref(lag), C.byref(ch), C.byref(val))
    return bool(found), lag.value, ch.value, val.value


def xcorr_at(y, x, lag: int) -> float:
    a = np.ascontiguousarray(y, dtype=np.float32).ravel()
    s = np.ascontiguousarray(x, dtype=np.float32).ravel()
    return lib().orc_xcorr_at(_p(a), a.size, _p(s), s.size, lag)


# ------------------------------------------------------------------ format convert
def pcm_to_planar(raw: np.ndarray, fmt: int, src_ch: int, dst_ch: int | None = None) -> np.ndarray:
    raw = np.ascontiguousarray(raw).view(np.uint8).ravel()
    frames = raw.size // (_BYTES[fmt] * src_ch)
    dst_ch = src_ch if dst_ch is None else dst_ch
    out = np.empty((dst_ch, frames), dtype=np.float32)
    lib().orc_pcm_to_planar(raw.ctypes.data, fmt, src_ch, frames, _chan_ptrs(out), dst_ch)
    return out


def planar_to_pcm24(buf) -> np.ndarray:
    a = _planar(buf)
    out = np.empty(a.shape[0] * a.shape[1] * 3, dtype=np.uint8)
    lib().orc_planar_to_pcm24(_chan_ptrs(a), a.shape[0], a.shape[1], out.ctypes.data)
    return out


def interleave(buf) -> np.ndarray:
    a = _planar(buf)
    out = np.empty(a.shape[0] * a.shape[1], dtype=np.float32)
    lib().orc_interleave(_chan_ptrs(a), a.shape[0], a.shape[1], _p(out))
    return out


def deinterleave(audio, num_ch: int) -> np.ndarray:
    a = np.ascontiguousarray(audio, dtype=np.float32).ravel()
    frames = a.size // num_ch
    out = np.empty((num_ch, frames), dtype=np.float32)
    lib().orc_deinterleave(_p(a), num_ch, frames, _chan_ptrs(out))
    return out


def resampling_source_coeffs(ratio: float) -> np.ndarray:
    c = (C.c_double * 6)()
    lib().orc_resampling_source_coeffs(ratio, c)
    return np.array(list(c))


class ResamplingAudioSource:
    """juce::ResamplingAudioSource restated (f9_oracle.cpp): the input source is a planar array read in order."""

    def __init__(self, source, intel_flush: bool = True):
        self._src = _planar(source)
        self._cursor = 0
        self._h = lib().orc_ras_create(self._src.shape[0], int(intel_flush))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_ras_destroy(self._h)
            self._h = None

    def set_resampling_ratio(self, ratio: float):
        lib().orc_ras_set_ratio(self._h, ratio)

    def prepare_to_play(self, samples_per_block: int):
        lib().orc_ras_prepare(self._h, samples_per_block)

    def flush_buffers(self):
        lib().orc_ras_flush(self._h)

    def get_next_audio_block(self, num_samples: int) -> np.ndarray:
        out = np.empty((self._src.shape[0], num_samples), dtype=np.float32)
        self.pulled = lib().orc_ras_get_next_block(self._h, _chan_ptrs(self._src), self._src.shape[1], self._cursor, _chan_ptrs(out), num_samples)
        self._cursor += self.pulled
        return out


def ras_convert(buf, ratio: float, num_out: int, block: int = 512, intel_flush: bool = True) -> np.ndarray:
    a = _planar(buf)
    out = np.empty((a.shape[0], num_out), dtype=np.float32)
    lib().orc_ras_convert(a.shape[0], _chan_ptrs(a), a.shape[1], ratio, _chan_ptrs(out), num_out, block, int(intel_flush))
    return out
