"""Synthetic workloads of BASELINE.json's configs (shapes from SURVEY.md 8(d)).

Signals: the stimuli of the reference -- 1 kHz sine at amplitude 0.5 (Source/MainComponent.cpp:907-932) with an exponential decay
into a -96 dBFS noise floor, delayed by a per-file round-trip latency; impulses of 0.9 (Source/MainComponent.cpp:934-945); sweeps
for the long multichannel file; seeds = file index.  torch is used here only to fill device / pinned host memory.
Sharding lives in the product (f9_multi_partition / f9_shard_units, include/f9dsp.h section H), not here.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

CONFIGS = {
    # name: (fs_in, fs_out, channels, source seconds, files)
    "config1_60s_stereo_44k1_to_48k": (44100, 48000, 2, 60.0, 1),
    "config2_256x_stereo_96k_to_44k1_trim_tail": (96000, 44100, 2, 10.0, 256),
    "config3_64ch_48k_to_192k_10min": (48000, 192000, 64, 600.0, 1),
    "config5_4096_mixed_to_48k": (0, 48000, 2, 10.0, 4096),          # fs_in round-robin over MIXED_RATES
}
MIXED_RATES = (44100, 48000, 88200, 96000, 192000)
DEFAULT = "config2_256x_stereo_96k_to_44k1_trim_tail"
TAIL_SECONDS = 0.5          # capture continues this long after source + latency (reverb-mode style capture)


@dataclass
class FileSpec:
    index: int                # global file index (seed)
    fs_in: int
    fs_out: int
    num_ch: int
    src_frames: int
    cap_frames: int           # frames per channel captured (padded to 64)
    latency: int              # round-trip latency in frames
    tail: bool                # reverb-tail scan requested
    shape: str                # "burst" (sine burst + decay + noise floor) | "sweep" (per-channel log sweep)


@dataclass
class Batch:
    name: str
    fs_in: int
    fs_out: int
    num_ch: int
    src_frames: int
    cap_frames: int           # frames per channel actually captured (same for every file; padded to 64)
    files: int
    latency_frames: list      # per file
    specs: list = field(default_factory=list)


def latency_of(file_index: int) -> int:
    return 128 * (file_index % 256) + 7            # L_i = 128*k + 7 frames


def file_specs(name: str, files: int | None = None, first_file: int = 0, seconds: float | None = None) -> list[FileSpec]:
    fs_in, fs_out, ch, secs, nfiles = CONFIGS[name]
    files = nfiles if files is None else files
    secs = secs if seconds is None else seconds
    out = []
    for k in range(files):
        i = first_file + k
        fi = MIXED_RATES[i % len(MIXED_RATES)] if fs_in == 0 else fs_in
        src = int(round(secs * fi))
        if name.startswith("config2") or name.startswith("config5"):
            lat = latency_of(i)
            cap = src + max(latency_of(j) for j in range(256)) + int(TAIL_SECONDS * fi)
            tail = name.startswith("config2")
            shape = "burst"
        else:
            lat, cap, tail, shape = 0, src, False, ("sweep" if name.startswith("config3") else "burst")
        out.append(FileSpec(i, fi, fs_out, ch, src, (cap + 63) // 64 * 64, lat, tail, shape))
    return out


def describe(name: str, files: int | None = None, first_file: int = 0) -> Batch:
    """Uniform-rate configs as one record (the CPU baseline leg and the config-2 tests use this form)."""
    specs = file_specs(name, files, first_file)
    s0 = specs[0]
    return Batch(name, s0.fs_in, s0.fs_out, s0.num_ch, s0.src_frames, s0.cap_frames, len(specs), [s.latency for s in specs], specs)


def _noise(idx, seed: int, xp):
    """Position-addressable uniform noise in [-1, 1): an integer hash of (seed, sample index), so any window of any channel can be
    generated on its own (time segments on different GPUs see the same signal)."""
    h = ((idx.astype(xp.int64) if xp is np else idx.to(xp.int64)) * 2654435761 + (seed * 40503 + 12345)) & 0xFFFFFFFF
    h = ((h ^ (h >> 16)) * 0x45d9f3b) & 0xFFFFFFFF           # products stay below 2^59: no int64 wrap on either backend
    h = ((h ^ (h >> 16)) * 0x45d9f3b) & 0xFFFFFFFF
    h = h ^ (h >> 16)
    return (h & 0xFFFFFF).astype(np.float64) / float(1 << 23) - 1.0 if xp is np else (h & 0xFFFFFF).to(xp.float64) / float(1 << 23) - 1.0


def window(spec: FileSpec, ch0: int, num_ch: int, first: int, frames: int, device=None):
    """float32 [num_ch, frames]: samples [first, first + frames) of channels [ch0, ch0 + num_ch) of the file's capture (numpy
    when device is None, else a CUDA tensor).  Frames outside the capture are zero."""
    if device is None:
        xp = np
        idx = np.arange(first, first + frames, dtype=np.int64)
    else:
        import torch as xp
        idx = xp.arange(first, first + frames, device=device, dtype=xp.int64)
    t = (idx - spec.latency).astype(np.float64) / spec.fs_in if xp is np else (idx - spec.latency).to(xp.float64) / spec.fs_in
    src_t = spec.src_frames / spec.fs_in
    rows = []
    valid = (idx >= 0) & (idx < spec.cap_frames)
    if spec.shape != "sweep":
        burst = 0.6 * src_t
        env = xp.where(t < burst, xp.ones_like(t), xp.exp(-(t - burst) * (14.0 / max(src_t - burst, 1e-3))))
        env = xp.where((t >= 0) & (t < src_t) & valid, env, xp.zeros_like(env))
        common = 0.5 * xp.sin(2 * np.pi * 1000.0 * t) * env
    for c in range(ch0, ch0 + num_ch):
        if spec.shape == "sweep":
            # log sweep 20 Hz -> 20 kHz over the file, channel-dependent start phase, amplitude 0.5
            T = src_t
            k = np.log(1000.0)
            ph = 2 * np.pi * 20.0 * T / k * ((xp.exp(t / T * k)) - 1.0) + 0.37 * c
            sig = xp.where((t >= 0) & (t < src_t) & valid, 0.5 * xp.sin(ph), xp.zeros_like(t))
        else:
            gain = 1.0 - 0.2 * c / max(spec.num_ch, 1)
            sig = gain * common + xp.where(valid, _noise(idx, spec.index * 131 + c, xp) * (10 ** (-96 / 20)), xp.zeros_like(t))
        rows.append(sig.astype(np.float32) if xp is np else sig.to(xp.float32))
    return np.stack(rows) if xp is np else xp.stack(rows)


def fill_device(batch: Batch, first_file: int, device, chunk_files: int = 32):
    """float32 CUDA tensor [files, ch, cap_frames] with the synthetic captures of a uniform batch."""
    import torch

    out = torch.empty((batch.files, batch.num_ch, batch.cap_frames), dtype=torch.float32, device=device)
    for i, spec in enumerate(batch.specs):
        out[i] = window(spec, 0, spec.num_ch, 0, batch.cap_frames, device)
    return out


def fill_host_numpy(batch: Batch, first_file: int, files: int) -> np.ndarray:
    """The same captures on the host (CPU baseline sample; no GPU needed)."""
    out = np.empty((files, batch.num_ch, batch.cap_frames), np.float32)
    specs = file_specs(batch.name, files, first_file)
    for i, spec in enumerate(specs):
        out[i] = window(spec, 0, spec.num_ch, 0, batch.cap_frames)
    return out


def float_to_s24(planes: np.ndarray) -> np.ndarray:
    """[ch, frames] float -> interleaved little-endian 24-bit payload (numpy; a 24-bit WAV file's data chunk)."""
    q = np.clip(np.round(planes.T.astype(np.float64) * 8388607.0), -8388608, 8388607).astype(np.int64)
    return np.stack([(q >> (8 * b)) & 0xff for b in range(3)], axis=-1).astype(np.uint8).ravel()
