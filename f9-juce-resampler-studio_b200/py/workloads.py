"""Synthetic workloads of BASELINE.json's configs (shapes from SURVEY.md 8(d)) and the sharding rules of 8(e).

Signals: 1 kHz sine at amplitude 0.5 (Source/MainComponent.cpp:149) with an exponential decay into a -96 dBFS noise
floor, delayed by a per-file round-trip latency; impulses of 0.9 (Source/MainComponent.cpp:938); seeds = file index.
torch is used here only to fill device / pinned host memory.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

CONFIGS = {
    # name: (fs_in, fs_out, channels, source seconds, files)
    "config1_60s_stereo_44k1_to_48k": (44100, 48000, 2, 60.0, 1),
    "config2_256x_stereo_96k_to_44k1_trim_tail": (96000, 44100, 2, 10.0, 256),
    "config3_64ch_48k_to_192k_10min": (48000, 192000, 64, 600.0, 1),
}
DEFAULT = "config2_256x_stereo_96k_to_44k1_trim_tail"
TAIL_SECONDS = 0.5          # capture continues this long after source + latency (reverb-mode style capture)


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


def latency_of(file_index: int) -> int:
    return 128 * (file_index % 256) + 7            # L_i = 128*k + 7 frames


def describe(name: str, files: int | None = None, first_file: int = 0) -> Batch:
    fs_in, fs_out, ch, secs, nfiles = CONFIGS[name]
    files = nfiles if files is None else files
    src = int(round(secs * fs_in))
    lats = [latency_of(first_file + i) for i in range(files)]
    cap = src + max(latency_of(k) for k in range(256)) + int(TAIL_SECONDS * fs_in)
    cap = (cap + 63) // 64 * 64
    return Batch(name, fs_in, fs_out, ch, src, cap, files, lats)


def fill_device(batch: Batch, first_file: int, device, chunk_files: int = 32):
    """Returns a float32 CUDA tensor [files, ch, cap_frames] with the synthetic captures."""
    import torch

    out = torch.empty((batch.files, batch.num_ch, batch.cap_frames), dtype=torch.float32, device=device)
    t = torch.arange(batch.cap_frames, device=device, dtype=torch.float64) / batch.fs_in
    src_t = batch.src_frames / batch.fs_in
    for f0 in range(0, batch.files, chunk_files):
        f1 = min(batch.files, f0 + chunk_files)
        lat = torch.tensor(batch.latency_frames[f0:f1], device=device, dtype=torch.float64)[:, None] / batch.fs_in
        tt = t[None, :] - lat                                             # time since the source started
        burst = 0.6 * src_t
        env = torch.where(tt < burst, torch.ones_like(tt), torch.exp(-(tt - burst) * (14.0 / max(src_t - burst, 1e-3))))
        env = torch.where((tt >= 0) & (tt < src_t), env, torch.zeros_like(env))
        sig = (0.5 * torch.sin(2 * np.pi * 1000.0 * tt) * env).to(torch.float32)
        for i in range(f0, f1):
            g = torch.Generator(device=device)
            g.manual_seed(first_file + i)
            noise = torch.randn((batch.num_ch, batch.cap_frames), generator=g, device=device, dtype=torch.float32) * (10 ** (-96 / 20))
            gains = 1.0 - 0.2 * torch.arange(batch.num_ch, device=device, dtype=torch.float32)[:, None] / max(batch.num_ch, 1)
            out[i] = sig[i - f0][None, :] * gains + noise
    return out


def fill_host_numpy(batch: Batch, first_file: int, files: int) -> np.ndarray:
    """CPU generation of the same shape of signal for the CPU baseline sample (no GPU needed)."""
    out = np.empty((files, batch.num_ch, batch.cap_frames), np.float32)
    t = np.arange(batch.cap_frames, dtype=np.float64) / batch.fs_in
    src_t = batch.src_frames / batch.fs_in
    burst = 0.6 * src_t
    for i in range(files):
        tt = t - latency_of(first_file + i) / batch.fs_in
        env = np.where(tt < burst, 1.0, np.exp(-(tt - burst) * (14.0 / max(src_t - burst, 1e-3))))
        env = np.where((tt >= 0) & (tt < src_t), env, 0.0)
        sig = (0.5 * np.sin(2 * np.pi * 1000.0 * tt) * env).astype(np.float32)
        rng = np.random.default_rng(first_file + i)
        for c in range(batch.num_ch):
            out[i, c] = sig * np.float32(1.0 - 0.2 * c / max(batch.num_ch, 1)) + (rng.standard_normal(batch.cap_frames) * 10 ** (-96 / 20)).astype(np.float32)
    return out


# ------------------------------------------------------------------ sharding: no collective, results gathered on the host
def shard_units(costs: list[int], world: int) -> list[list[int]]:
    """Greedy bin-packing of work units (file, channel or segment) by output-sample count onto `world` GPUs."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    loads = [0] * world
    bins: list[list[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (loads[k], k))
        bins[r].append(i)
        loads[r] += costs[i]
    return [sorted(b) for b in bins]


def time_segments(num_out: int, seg_out: int) -> list[tuple[int, int]]:
    """Split outputs [0, num_out) of one long channel into (n0, count) segments of at most seg_out outputs."""
    return [(n0, min(seg_out, num_out - n0)) for n0 in range(0, num_out, seg_out)]
