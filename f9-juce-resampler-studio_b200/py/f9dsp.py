"""ctypes binding of libf9dsp.so (include/f9dsp.h) with numpy front-ends.

This is test / benchmark plumbing over the C ABI: every call below goes straight into the
CUDA library.  There is no fallback of any kind: a missing library raises ImportError-like
RuntimeError at load, a missing GPU raises F9Error at context creation.
Names mirror the reference's helpers (Source/MainComponent.h:197-237, Source/AppState.h:221-258).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "lib", "libf9dsp.so")

OK, ERR_INVALID, ERR_CUDA, ERR_NOMEM, ERR_NO_DEVICE, ERR_UNSUPPORTED = 0, -1, -2, -3, -4, -5
WINDOWED_SINC, LAGRANGE, CATMULL_ROM, LINEAR, ZERO_ORDER_HOLD = 0, 1, 2, 3, 4
PCM_U8, PCM_S16LE, PCM_S24LE, PCM_S32LE, PCM_F32LE = 1, 2, 3, 4, 5
TAIL_RMS, TAIL_PEAK = 0, 1
JOB_TAIL_SCAN, JOB_REMOVE_DC, JOB_PCM24, JOB_DC_REFERENCE_ORDER = 1, 2, 4, 8
_BYTES = {PCM_U8: 1, PCM_S16LE: 2, PCM_S24LE: 3, PCM_S32LE: 4, PCM_F32LE: 4}

_fp = C.POINTER(C.c_float)
_fpp = C.POINTER(_fp)


class F9Error(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"f9dsp error {code}: {msg}")
        self.code = code


class DevBuffer(C.Structure):
    _fields_ = [("base", C.c_void_p), ("ch_stride", C.c_longlong), ("numCh", C.c_int), ("numFrames", C.c_int)]


class ResampleSeg(C.Structure):
    _fields_ = [("inp", C.c_void_p), ("in_offset", C.c_longlong), ("in_avail", C.c_longlong),
                ("out", C.c_void_p), ("n0", C.c_longlong), ("num_out", C.c_longlong)]


class TailParams(C.Structure):
    _fields_ = [("start_frame", C.c_longlong), ("window", C.c_int), ("hop", C.c_int), ("required", C.c_int),
                ("mode", C.c_int), ("has_nf", C.c_int), ("nf_db", C.c_float), ("margin_pct", C.c_float)]


class Job(C.Structure):
    _fields_ = [("captured", _fpp), ("numCh", C.c_int), ("captured_frames", C.c_int), ("latency_samples", C.c_int),
                ("original_length", C.c_int), ("fs_in", C.c_double), ("fs_out", C.c_double), ("interp_kind", C.c_int),
                ("flags", C.c_int), ("tail_window", C.c_int), ("tail_hop", C.c_int), ("tail_required", C.c_int),
                ("tail_mode", C.c_int), ("has_nf", C.c_int), ("nf_db", C.c_float), ("margin_pct", C.c_float),
                ("out", _fpp), ("out_capacity", C.c_int), ("out_pcm24", C.c_void_p),
                ("src_pcm", C.c_void_p), ("src_fmt", C.c_int), ("src_ch", C.c_int)]


class Unit(C.Structure):
    _fields_ = [("job", C.c_int), ("device", C.c_int), ("ch0", C.c_int), ("num_ch", C.c_int), ("n0", C.c_longlong),
                ("num_out", C.c_longlong), ("tail_only", C.c_int), ("reserved", C.c_int), ("cost", C.c_longlong)]


class Result(C.Structure):
    _fields_ = [("status", C.c_int), ("latency_frames", C.c_int), ("trim_start", C.c_int), ("frames_copied", C.c_int),
                ("out_frames", C.c_int), ("tail_stop_frame", C.c_longlong), ("tail_polls", C.c_int)]


# every symbol include/f9dsp.h declares: (name, restype, argtypes)
_vp, _i, _ll, _f, _d = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_double
_ip, _llp, _dp = C.POINTER(C.c_int), C.POINTER(C.c_longlong), C.POINTER(C.c_double)
SYMBOLS = [
    ("f9_context_create", _i, [_i, C.POINTER(_vp)]),
    ("f9_context_destroy", None, [_vp]),
    ("f9_last_error", C.c_char_p, [_vp]),
    ("f9_set_stream", _i, [_vp, _vp]),
    ("f9_synchronize", _i, [_vp]),
    ("f9_context_set_option", _i, [_vp, C.c_char_p, _i]),
    ("f9_context_clear_options", _i, [_vp]),
    ("f9_launch_count", _ll, [_vp]),
    ("f9_host_alloc", _i, [_vp, C.POINTER(_vp), C.c_size_t]),
    ("f9_host_free", _i, [_vp, _vp]),
    ("f9_version", _i, []),
    ("f9_device_count", _i, []),
    ("f9_umma_selfcheck", _d, [_i, _ll, _ll, _ip]),
    ("f9_hankel_selfcheck", _d, [_i, _i, _ip]),
    ("f9_recording_length", _i, [_i, _i]),
    ("f9_noise_floor_threshold_db", _f, [_i, _f, _f]),
    ("f9_threshold_linear", _f, [_f]),
    ("f9_latency_ms", _d, [_i, _d]),
    ("f9_needs_latency_remeasurement", _i, [_i, _i, _i]),
    ("f9_find_peak_position", _i, [_vp, _fpp, _i, _i, _f, _ip]),
    ("f9_find_peak_interleaved", _i, [_vp, _fp, _ll, _f, _llp, _ip]),
    ("f9_calculate_rms", _i, [_vp, _fpp, _i, _i, _fp]),
    ("f9_calculate_noise_floor_db", _i, [_vp, _fpp, _i, _i, _fp]),
    ("f9_is_reverb_tail_below_noise_floor", _i, [_vp, _fpp, _i, _i, _i, _f, _f, _ip]),
    ("f9_is_reverb_tail_below_noise_floor_swift", _i, [_vp, _fp, _ll, _i, _f, _f, _ip]),
    ("f9_trim_latency", _i, [_vp, _fpp, _i, _i, _i, _i, _fpp, _ip]),
    ("f9_trim_latency_swift", _i, [_vp, _fp, _ll, _ll, _ll, _i, _fp, _llp]),
    ("f9_remove_dc_offset", _i, [_vp, _fpp, _i, _i]),
    ("f9_remove_dc_offset_ex", _i, [_vp, _fpp, _i, _i, _i]),
    ("f9_generate_impulse", _i, [_vp, _fpp, _i, _i]),
    ("f9_generate_sine_wave", _i, [_vp, _fpp, _i, _i, _f, _f, _f, _fp, _i]),
    ("f9_generate_sine_wave_swift", _i, [_vp, _fp, _i, _i, _d, _d, _f, _dp]),
    ("f9_dev_generate_impulse", _i, [_vp, C.POINTER(DevBuffer), _i]),
    ("f9_dev_generate_sine_wave", _i, [_vp, C.POINTER(DevBuffer), _f, _f, _f, _f]),
    ("f9_tail_scan", _i, [_vp, _fpp, _i, _ll, _ll, _i, _i, _i, _i, _i, _f, _f, _llp, _ip, _i, _ip]),
    ("f9_xcorr_peak", _i, [_vp, _fpp, _i, _i, _fp, _i, _i, _i, _f, _ip, _ip, _ip, _dp]),
    ("f9_interp_create", _i, [_vp, _i, C.POINTER(_vp)]),
    ("f9_interp_destroy", None, [_vp]),
    ("f9_interp_reset", _i, [_vp]),
    ("f9_interp_base_latency", _f, [_vp]),
    ("f9_interp_process", _i, [_vp, _d, _fp, _fp, _i]),
    ("f9_interp_process_adding", _i, [_vp, _d, _fp, _fp, _i, _f]),
    ("f9_interp_process_wrap", _i, [_vp, _d, _fp, _fp, _i, _i, _i]),
    ("f9_ras_create", _i, [_vp, _i, C.POINTER(_vp)]),
    ("f9_ras_destroy", None, [_vp]),
    ("f9_ras_set_resampling_ratio", _i, [_vp, _d]),
    ("f9_ras_get_resampling_ratio", _d, [_vp]),
    ("f9_ras_prepare_to_play", _i, [_vp, _i, _d]),
    ("f9_ras_flush_buffers", _i, [_vp]),
    ("f9_ras_release_resources", _i, [_vp]),
    ("f9_ras_set_intel_denormal_flush", _i, [_vp, _i]),
    ("f9_ras_num_samples_to_pull", _i, [_vp, _i]),
    ("f9_ras_get_next_audio_block", _i, [_vp, _fpp, _i, _fpp, _i]),
    ("f9_ras_convert", _i, [_vp, _fpp, _i, _ll, _d, _fpp, _ll]),
    ("f9_ras_scratch_frames", _ll, [_d, _ll]),
    ("f9_dev_ras_convert", _i, [_vp, _vp, _ll, _i, _ll, _d, _vp, _ll, _ll, _vp, _ll]),
    ("f9_sinc_table_set", _i, [_vp, _fp]),
    ("f9_sinc_table_get", _i, [_vp, _fp]),
    ("f9_resampled_length", _ll, [_ll, _d, _d]),
    ("f9_process_batch", _i, [_vp, C.POINTER(Job), _i, C.POINTER(Result)]),
    ("f9_multi_create", _i, [_ip, _i, C.POINTER(_vp)]),
    ("f9_multi_destroy", None, [_vp]),
    ("f9_multi_device_count", _i, [_vp]),
    ("f9_multi_context", _vp, [_vp, _i]),
    ("f9_multi_last_error", C.c_char_p, [_vp]),
    ("f9_multi_process_batch", _i, [_vp, C.POINTER(Job), _i, C.POINTER(Result), _ip]),
    ("f9_shard_units", _i, [_llp, _i, _i, _ip]),
    ("f9_multi_partition", _i, [C.POINTER(Job), _i, _i, _ll, C.POINTER(Unit), _i, _ip]),
    ("f9_process_units", _i, [_vp, C.POINTER(Job), _i, C.POINTER(Unit), _i, _i, C.POINTER(Result)]),
    ("f9_merge_unit_results", _i, [C.POINTER(Job), _i, C.POINTER(Unit), C.POINTER(Result), _i, C.POINTER(Result)]),
    ("f9_dev_find_peak_batch", _i, [_vp, C.POINTER(DevBuffer), _i, _f, _vp]),
    ("f9_measure_latency", _i, [_vp, _fpp, _i, _i, C.c_float, _ip, _fp]),
    ("f9_dev_latency_stats_batch", _i, [_vp, C.POINTER(DevBuffer), _i, C.c_float, _vp, _vp, _vp]),
    ("f9_dev_stats_batch", _i, [_vp, C.POINTER(DevBuffer), _i, _vp, _vp]),
    ("f9_resample_plan_create", _i, [_vp, _i, _d, C.POINTER(ResampleSeg), _i, C.POINTER(_vp)]),
    ("f9_resample_plan_run", _i, [_vp]),
    ("f9_plan_destroy", None, [_vp]),
    ("f9_resample_segment_input_range", _i, [_i, _d, _ll, _ll, _llp, _llp]),
    ("f9_dev_tail_scan_batch", _i, [_vp, C.POINTER(DevBuffer), C.POINTER(TailParams), _i, _vp, _vp, _i]),
    ("f9_dev_trim_batch", _i, [_vp, C.POINTER(DevBuffer), _ip, C.POINTER(DevBuffer), _i, _i]),
    ("f9_dev_xcorr_peak_batch", _i, [_vp, C.POINTER(DevBuffer), _i, _vp, _i, _i, _i, _vp]),
    ("f9_pcm_to_planar", _i, [_vp, _vp, _i, _i, _ll, _fpp, _i]),
    ("f9_planar_to_pcm24", _i, [_vp, _fpp, _i, _ll, _vp]),
    ("f9_interleave", _i, [_vp, _fpp, _i, _ll, _fp]),
    ("f9_deinterleave", _i, [_vp, _fp, _i, _ll, _fpp]),
    ("f9_dev_pcm_to_planar", _i, [_vp, _vp, _i, _i, _ll, _vp, _ll, _i]),
    ("f9_dev_planar_to_pcm24", _i, [_vp, _vp, _ll, _i, _ll, _vp]),
    ("f9_dev_pcm_to_planar_batch", _i, [_vp, C.POINTER(C.c_void_p), _i, _i, C.POINTER(DevBuffer), _i]),
    ("f9_dev_planar_to_pcm24_batch", _i, [_vp, C.POINTER(DevBuffer), C.POINTER(C.c_void_p), _i]),
]

_lib = None


def lib():
    """Load libf9dsp.so.  Raises RuntimeError when it has not been built -- never falls back."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(or `make -C f9-juce-resampler-studio_b200`). There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, res, args in SYMBOLS:
            fn = getattr(L, name)           # AttributeError if the header and the library disagree
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def device_count() -> int:
    return lib().f9_device_count()


# ------------------------------------------------------------------ host scalars
def recording_length(src: int, lat: int) -> int:
    return lib().f9_recording_length(src, lat)


def noise_floor_threshold_db(has_nf: bool, nf_db: float, margin: float) -> np.float32:
    return np.float32(lib().f9_noise_floor_threshold_db(int(has_nf), nf_db, margin))


def threshold_linear(db: float) -> np.float32:
    return np.float32(lib().f9_threshold_linear(db))


def latency_ms(samples: int, fs: float) -> float:
    return lib().f9_latency_ms(samples, fs)


def needs_latency_remeasurement(measured: int, last_buf: int, buf: int) -> bool:
    return bool(lib().f9_needs_latency_remeasurement(measured, last_buf, buf))


def resampled_length(n_in: int, fs_in: float, fs_out: float) -> int:
    return lib().f9_resampled_length(n_in, fs_in, fs_out)


def segment_input_range(kind: int, ratio: float, n0: int, num_out: int):
    a, b = C.c_longlong(0), C.c_longlong(0)
    rc = lib().f9_resample_segment_input_range(kind, ratio, n0, num_out, C.byref(a), C.byref(b))
    if rc:
        raise F9Error(rc, "bad segment")
    return a.value, b.value


def default_sinc_table() -> np.ndarray:
    t = np.empty(10001, dtype=np.float32)
    lib().f9_sinc_table_get(None, t.ctypes.data_as(_fp))
    return t


# ------------------------------------------------------------------ helpers
def _planar(a) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float32)
    if a.ndim == 1:
        a = a[None, :]
    assert a.ndim == 2
    return a


def _chan_ptrs(a: np.ndarray):
    n = a.shape[0]
    arr = (_fp * max(n, 1))()
    for c in range(n):
        arr[c] = C.cast(a.ctypes.data + c * a.strides[0], _fp)
    return arr


def _p(a: np.ndarray):
    return a.ctypes.data_as(_fp)


class BuiltJobs:
    """f9_job array built from job dicts, with everything it points to kept alive."""

    def __init__(self, n):
        self.n = n
        self.J = (Job * max(n, 1))()
        self.R = (Result * max(n, 1))()
        self.keep, self.outs, self.pcms = [], [], []

    def finish(self, rc, check):
        R, n = self.R, self.n
        res = [dict(status=R[i].status, latency_frames=R[i].latency_frames, trim_start=R[i].trim_start,
                    frames_copied=R[i].frames_copied, out_frames=R[i].out_frames,
                    tail_stop_frame=R[i].tail_stop_frame, tail_polls=R[i].tail_polls) for i in range(n)]
        if rc < 0 and all(r["status"] == 0 for r in res):
            check(rc)
        outputs = [o[:, :f] for (o, f) in self.outs]
        pcm_out = [None if p is None else p[: res[i]["out_frames"] * outputs[i].shape[0] * 3] for i, p in enumerate(self.pcms)]
        return outputs, pcm_out, res


def build_jobs(jobs: list[dict]) -> BuiltJobs:
    B = BuiltJobs(len(jobs))
    J, keep = B.J, B.keep
    for i, j in enumerate(jobs):
        fs_in, fs_out = float(j.get("fs_in", 44100.0)), float(j.get("fs_out", 44100.0))
        out_frames = resampled_length(j["original_length"], fs_in, fs_out) if fs_in != fs_out else j["original_length"]
        if "src_pcm" in j:          # the capture as file bytes: (raw uint8 array, fmt, src_ch[, numCh])
            raw, fmt, src_ch = j["src_pcm"][:3]
            raw = np.ascontiguousarray(raw).view(np.uint8).ravel()
            num_ch = j["src_pcm"][3] if len(j["src_pcm"]) > 3 else src_ch
            frames = raw.size // (_BYTES[fmt] * src_ch)
            J[i].src_pcm, J[i].src_fmt, J[i].src_ch = raw.ctypes.data, fmt, src_ch
            J[i].numCh, J[i].captured_frames = num_ch, frames
            keep.append(raw)
        else:
            cap = _planar(j["captured"])
            cp = _chan_ptrs(cap)
            keep += [cap, cp]
            num_ch = cap.shape[0]
            J[i].captured, J[i].numCh, J[i].captured_frames = cp, cap.shape[0], cap.shape[1]
        flags = 0
        if j.get("no_float_out"):
            out = np.zeros((num_ch, 0), dtype=np.float32)
        else:
            out = np.full((num_ch, max(out_frames, 1)), np.nan, dtype=np.float32)
            op = _chan_ptrs(out)
            keep += [out, op]
            J[i].out, J[i].out_capacity = op, out.shape[1]
        J[i].latency_samples, J[i].original_length = j["latency_samples"], j["original_length"]
        J[i].fs_in, J[i].fs_out, J[i].interp_kind = fs_in, fs_out, j.get("kind", WINDOWED_SINC)
        if "tail" in j:
            w, h, r, m, has, nf, mg = j["tail"]
            flags |= JOB_TAIL_SCAN
            J[i].tail_window, J[i].tail_hop, J[i].tail_required, J[i].tail_mode = w, h, r, m
            J[i].has_nf, J[i].nf_db, J[i].margin_pct = int(has), nf, mg
        if j.get("remove_dc"):
            flags |= JOB_REMOVE_DC
            if j.get("remove_dc") == "reference":
                flags |= JOB_DC_REFERENCE_ORDER
        pcm = None
        if j.get("pcm24"):
            flags |= JOB_PCM24
            pcm = np.zeros(max(out_frames, 1) * num_ch * 3, dtype=np.uint8)
            J[i].out_pcm24 = pcm.ctypes.data
            keep.append(pcm)
        J[i].flags = flags
        B.outs.append((out, out_frames if not j.get("no_float_out") else 0))
        B.pcms.append(pcm)
    return B


class Multi:
    """f9_multi: the batch job flow over several GPUs (one context and one host thread per device-list entry)."""

    def __init__(self, devices):
        devs = (C.c_int * len(devices))(*devices)
        self._h = C.c_void_p(None)
        rc = lib().f9_multi_create(devs, len(devices), C.byref(self._h))
        if rc:
            raise F9Error(rc, (lib().f9_last_error(None) or b"").decode())
        self.devices = list(devices)

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().f9_multi_destroy(self._h)
            self._h = C.c_void_p(None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc < 0:
            raise F9Error(rc, (lib().f9_multi_last_error(self._h) or b"").decode())
        return rc

    def process_batch(self, jobs: list[dict]):
        B = build_jobs(jobs)
        dev = (C.c_int * max(B.n, 1))()
        rc = lib().f9_multi_process_batch(self._h, B.J, B.n, B.R, dev)
        outs, pcms, res = B.finish(rc, self._check)
        for i, r in enumerate(res):
            r["device"] = dev[i]
        return outs, pcms, res


def partition(J, n_jobs: int, n_devices: int, seg_out: int = 0):
    """f9_multi_partition over an f9_job array: list of Unit."""
    cap = n_jobs + 64
    while True:
        U = (Unit * cap)()
        nu = C.c_int(0)
        rc = lib().f9_multi_partition(J, n_jobs, n_devices, seg_out, U, cap, C.byref(nu))
        if rc == ERR_NOMEM:
            cap = nu.value
            continue
        if rc:
            raise F9Error(rc, "partition failed")
        return U, nu.value


class Context:
    """One per host thread / GPU (f9_context)."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p(None)
        rc = lib().f9_context_create(device, C.byref(self._h))
        if rc:
            msg = lib().f9_last_error(None)
            raise F9Error(rc, (msg or b"").decode())
        self.device = device

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().f9_context_destroy(self._h)
            self._h = C.c_void_p(None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc: int):
        if rc < 0:
            raise F9Error(rc, (lib().f9_last_error(self._h) or b"").decode())
        return rc

    @property
    def handle(self):
        return self._h

    def set_stream(self, cuda_stream: int | None):
        self._check(lib().f9_set_stream(self._h, C.c_void_p(cuda_stream or 0)))

    def set_option(self, name: str, value: int = 1):
        """Variant switch for plans created afterwards (f9_context_set_option)."""
        self._check(lib().f9_context_set_option(self._h, name.encode(), int(value)))

    def clear_options(self):
        self._check(lib().f9_context_clear_options(self._h))

    def synchronize(self):
        self._check(lib().f9_synchronize(self._h))

    @property
    def launch_count(self) -> int:
        return lib().f9_launch_count(self._h)

    def pinned_empty(self, shape, dtype=np.float32) -> np.ndarray:
        """numpy array over pinned host memory (f9_host_alloc); kept alive by the context's list."""
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = C.c_void_p(None)
        self._check(lib().f9_host_alloc(self._h, C.byref(p), max(n, 1)))
        buf = (C.c_char * max(n, 1)).from_address(p.value)
        arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
        if not hasattr(self, "_pinned"):
            self._pinned = []
        self._pinned.append(p)
        return arr

    # ---- C. reference-shaped helpers
    def find_peak_position(self, buf, threshold: float) -> int:
        a = _planar(buf)
        out = C.c_int(0)
        self._check(lib().f9_find_peak_position(self._h, _chan_ptrs(a), a.shape[0], a.shape[1], threshold, C.byref(out)))
        return out.value

    def find_peak_interleaved(self, audio, threshold: float):
        a = np.ascontiguousarray(audio, dtype=np.float32).ravel()
        idx, found = C.c_longlong(0), C.c_int(0)
        self._check(lib().f9_find_peak_interleaved(self._h, _p(a), a.size, threshold, C.byref(idx), C.byref(found)))
        return idx.value, bool(found.value)

    def calculate_rms(self, buf) -> np.float32:
        a = _planar(buf)
        out = C.c_float(0)
        self._check(lib().f9_calculate_rms(self._h, _chan_ptrs(a), a.shape[0], a.shape[1], C.byref(out)))
        return np.float32(out.value)

    def calculate_noise_floor_db(self, buf) -> np.float32:
        a = _planar(buf)
        out = C.c_float(0)
        self._check(lib().f9_calculate_noise_floor_db(self._h, _chan_ptrs(a), a.shape[0], a.shape[1], C.byref(out)))
        return np.float32(out.value)

    def is_reverb_tail_below_noise_floor(self, window, has_nf: bool, nf_db: float, margin: float) -> bool:
        a = _planar(window)
        out = C.c_int(0)
        self._check(lib().f9_is_reverb_tail_below_noise_floor(self._h, _chan_ptrs(a), a.shape[0], a.shape[1],
                                                              int(has_nf), nf_db, margin, C.byref(out)))
        return bool(out.value)

    def is_reverb_tail_below_noise_floor_swift(self, window, has_nf: bool, nf_db: float, margin: float) -> bool:
        a = np.ascontiguousarray(window, dtype=np.float32).ravel()
        out = C.c_int(0)
        self._check(lib().f9_is_reverb_tail_below_noise_floor_swift(self._h, _p(a), a.size, int(has_nf), nf_db, margin, C.byref(out)))
        return bool(out.value)

    def trim_latency(self, captured, latency_samples: int, original_length: int):
        a = _planar(captured)
        out = np.empty((a.shape[0], max(original_length, 0)), dtype=np.float32)
        copied = C.c_int(0)
        self._check(lib().f9_trim_latency(self._h, _chan_ptrs(a), a.shape[0], a.shape[1], latency_samples, original_length,
                                          _chan_ptrs(out), C.byref(copied)))
        return out, copied.value

    def trim_latency_swift(self, captured, latency_samples: int, source_frames: int, channels: int) -> np.ndarray:
        a = np.ascontiguousarray(captured, dtype=np.float32).ravel()
        out = np.empty(max(source_frames * channels, 1), dtype=np.float32)
        n = C.c_longlong(0)
        self._check(lib().f9_trim_latency_swift(self._h, _p(a), a.size, latency_samples, source_frames, channels, _p(out), C.byref(n)))
        return out[: n.value].copy()

    def remove_dc_offset(self, buf, reference_order: bool = True) -> np.ndarray:
        a = _planar(buf).copy()
        self._check(lib().f9_remove_dc_offset_ex(self._h, _chan_ptrs(a), a.shape[0], a.shape[1], int(reference_order)))
        return a

    def generate_impulse(self, num_ch: int, num_frames: int) -> np.ndarray:
        out = np.full((num_ch, num_frames), np.nan, dtype=np.float32)
        self._check(lib().f9_generate_impulse(self._h, _chan_ptrs(out), num_ch, num_frames))
        return out

    def generate_sine_wave(self, num_ch: int, num_samples: int, phase: float = 0.0, frequency: float = 1000.0,
                           sample_rate: float = 44100.0, amplitude: float = 0.5, callback_form: bool = False):
        """MainComponent::generateSineWave: returns (buffer, new sinePhase)."""
        out = np.full((num_ch, num_samples), np.nan, dtype=np.float32)
        ph = C.c_float(phase)
        self._check(lib().f9_generate_sine_wave(self._h, _chan_ptrs(out), num_ch, num_samples, frequency, sample_rate, amplitude,
                                                C.byref(ph), int(callback_form)))
        return out, np.float32(ph.value)

    def generate_sine_wave_swift(self, frames: int, channels: int, phase: float = 0.0, frequency: float = 1000.0,
                                 sample_rate: float = 44100.0, amplitude: float = 0.5):
        out = np.full(max(frames * channels, 1), np.nan, dtype=np.float32)
        ph = C.c_double(phase)
        self._check(lib().f9_generate_sine_wave_swift(self._h, _p(out), frames, channels, frequency, sample_rate, amplitude, C.byref(ph)))
        return out[: frames * channels], ph.value

    def tail_scan(self, buf, start_frame: int, window: int, hop: int, required: int, mode: int,
                  has_nf: bool, nf_db: float, margin: float):
        a = _planar(buf)
        max_polls = max(0, (a.shape[1] - start_frame) // max(hop, 1)) + 1
        flags = np.full(max_polls, -2, dtype=np.int32)
        stop, polls = C.c_longlong(-1), C.c_int(0)
        self._check(lib().f9_tail_scan(self._h, _chan_ptrs(a), a.shape[0], a.shape[1], start_frame, window, hop, required, mode,
                                       int(has_nf), nf_db, margin, C.byref(stop), flags.ctypes.data_as(_ip), max_polls, C.byref(polls)))
        return stop.value, flags[: polls.value].copy()

    def xcorr_peak(self, y, x, lag_min: int, lag_max: int, threshold: float):
        a = _planar(y)
        s = np.ascontiguousarray(x, dtype=np.float32).ravel()
        found, lag, ch, val = C.c_int(0), C.c_int(0), C.c_int(0), C.c_double(0)
        self._check(lib().f9_xcorr_peak(self._h, _chan_ptrs(a), a.shape[0], a.shape[1], _p(s), s.size, lag_min, lag_max, threshold,
                                        C.byref(found), C.byref(lag), C.byref(ch), C.byref(val)))
        return bool(found.value), lag.value, ch.value, val.value

    # ---- D. interpolators
    def interpolator(self, kind: int) -> "Interpolator":
        return Interpolator(self, kind)

    def sinc_table_set(self, table):
        t = np.ascontiguousarray(table, dtype=np.float32)
        assert t.size == 10001
        self._check(lib().f9_sinc_table_set(self._h, _p(t)))

    def sinc_table_get(self) -> np.ndarray:
        t = np.empty(10001, dtype=np.float32)
        self._check(lib().f9_sinc_table_get(self._h, _p(t)))
        return t

    # ---- E. batch flow
    def process_batch(self, jobs: list[dict]):
        """jobs: dicts with captured (numCh x frames) or src_pcm=(raw bytes, fmt, src_ch[, numCh]), latency_samples, original_length,
        fs_in, fs_out, kind, and optional tail=(window, hop, required, mode, has_nf, nf_db, margin), remove_dc, pcm24, no_float_out.
        Returns (outputs, pcm, results)."""
        B = build_jobs(jobs)
        rc = lib().f9_process_batch(self._h, B.J, B.n, B.R)
        return B.finish(rc, self._check)

    # ---- G. format convert
    def pcm_to_planar(self, raw, fmt: int, src_ch: int, dst_ch: int | None = None) -> np.ndarray:
        raw = np.ascontiguousarray(raw).view(np.uint8).ravel()
        frames = raw.size // (_BYTES[fmt] * src_ch)
        dst_ch = src_ch if dst_ch is None else dst_ch
        out = np.empty((dst_ch, frames), dtype=np.float32)
        self._check(lib().f9_pcm_to_planar(self._h, raw.ctypes.data, fmt, src_ch, frames, _chan_ptrs(out), dst_ch))
        return out

    def planar_to_pcm24(self, buf) -> np.ndarray:
        a = _planar(buf)
        out = np.empty(a.shape[0] * a.shape[1] * 3, dtype=np.uint8)
        self._check(lib().f9_planar_to_pcm24(self._h, _chan_ptrs(a), a.shape[0], a.shape[1], out.ctypes.data))
        return out

    def interleave(self, buf) -> np.ndarray:
        a = _planar(buf)
        out = np.empty(a.shape[0] * a.shape[1], dtype=np.float32)
        self._check(lib().f9_interleave(self._h, _chan_ptrs(a), a.shape[0], a.shape[1], _p(out)))
        return out

    def deinterleave(self, audio, num_ch: int) -> np.ndarray:
        a = np.ascontiguousarray(audio, dtype=np.float32).ravel()
        frames = a.size // num_ch
        out = np.empty((num_ch, frames), dtype=np.float32)
        self._check(lib().f9_deinterleave(self._h, _p(a), num_ch, frames, _chan_ptrs(out)))
        return out

    # ---- juce::ResamplingAudioSource over whole channels from reset state
    def ras_convert(self, buf, ratio: float, num_out: int) -> np.ndarray:
        a = _planar(buf)
        out = np.empty((a.shape[0], num_out), dtype=np.float32)
        self._check(lib().f9_ras_convert(self._h, _chan_ptrs(a), a.shape[0], a.shape[1], ratio, _chan_ptrs(out), num_out))
        return out

    # ---- whole-channel conversion through a one-segment job (host buffers)
    def resample(self, buf, fs_in: float, fs_out: float, kind: int, num_out: int | None = None) -> np.ndarray:
        a = _planar(buf)
        outs, _, res = self.process_batch([dict(captured=a, latency_samples=0, original_length=a.shape[1],
                                                fs_in=fs_in, fs_out=fs_out, kind=kind)])
        if res[0]["status"]:
            raise F9Error(res[0]["status"], "resample failed")
        return outs[0] if num_out is None else outs[0][:, :num_out]


class ResamplingAudioSource:
    """juce::ResamplingAudioSource-shaped object over planar numpy input (the input AudioSource is `source`, read in order)."""

    def __init__(self, ctx: Context, source, num_channels: int | None = None):
        self._ctx = ctx
        self._src = _planar(source)
        self._cursor = 0
        self._nch = num_channels or self._src.shape[0]
        self._h = C.c_void_p(None)
        ctx._check(lib().f9_ras_create(ctx.handle, self._nch, C.byref(self._h)))

    def __del__(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().f9_ras_destroy(self._h)
            self._h = C.c_void_p(None)

    def set_resampling_ratio(self, ratio: float):
        self._ctx._check(lib().f9_ras_set_resampling_ratio(self._h, ratio))

    def prepare_to_play(self, samples_per_block: int, sample_rate: float = 44100.0):
        self._ctx._check(lib().f9_ras_prepare_to_play(self._h, samples_per_block, sample_rate))

    def flush_buffers(self):
        self._ctx._check(lib().f9_ras_flush_buffers(self._h))

    def get_next_audio_block(self, num_samples: int) -> np.ndarray:
        out = np.empty((self._nch, num_samples), dtype=np.float32)
        rest = np.ascontiguousarray(self._src[:, self._cursor:])
        pulled = self._ctx._check(lib().f9_ras_get_next_audio_block(self._h, _chan_ptrs(rest), rest.shape[1], _chan_ptrs(out), num_samples))
        self._cursor += min(pulled, rest.shape[1])
        self.pulled = pulled
        return out


class Interpolator:
    """juce::Interpolators::{WindowedSinc, Lagrange, ...}-shaped object: process(speedRatio, in, numOut)."""

    def __init__(self, ctx: Context, kind: int):
        self._ctx = ctx
        self._h = C.c_void_p(None)
        ctx._check(lib().f9_interp_create(ctx.handle, kind, C.byref(self._h)))

    def __del__(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().f9_interp_destroy(self._h)
            self._h = C.c_void_p(None)

    def reset(self):
        self._ctx._check(lib().f9_interp_reset(self._h))

    @property
    def base_latency(self) -> float:
        return lib().f9_interp_base_latency(self._h)

    def process(self, ratio: float, inp, num_out: int):
        a = np.ascontiguousarray(inp, dtype=np.float32)
        out = np.empty(num_out, dtype=np.float32)
        used = self._ctx._check(lib().f9_interp_process(self._h, ratio, _p(a), _p(out), num_out))
        return out, used

    def process_adding(self, ratio: float, inp, out: np.ndarray, gain: float) -> int:
        a = np.ascontiguousarray(inp, dtype=np.float32)
        assert out.dtype == np.float32 and out.flags.c_contiguous
        return self._ctx._check(lib().f9_interp_process_adding(self._h, ratio, _p(a), _p(out), out.size, gain))

    def process_wrap(self, ratio: float, inp, num_out: int, avail: int, wrap: int):
        a = np.ascontiguousarray(inp, dtype=np.float32)
        out = np.empty(num_out, dtype=np.float32)
        used = self._ctx._check(lib().f9_interp_process_wrap(self._h, ratio, _p(a), _p(out), num_out, avail, wrap))
        return out, used
