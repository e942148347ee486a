#!/usr/bin/env python
"""Interleaved A/B of the file-bytes e2e leg (config 2's 256 payloads): variants alternate step by step in one process, so that a
burst of slow steps hits all of them alike.  VARIANTS="name:kind:opt1=v,opt2;..." (kind 0 WindowedSinc, 1 Lagrange)."""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

f9 = g._load_pkg()
L = f9.lib()
files, ch, fs_in, fs_out, src = 256, 2, 96000, 44100, 960000
cap = (src + 128 * 255 + 7 + 48000 + 63) // 64 * 64
n_out = f9.resampled_length(src, fs_in, fs_out)
pay = torch.zeros(files * ch * cap * 3, dtype=torch.uint8, pin_memory=True)
out24 = torch.zeros(files * ch * n_out * 3, dtype=torch.uint8, pin_memory=True)


def jobs(kind):
    J = (f9.Job * files)()
    for i in range(files):
        j = J[i]
        lat = 128 * (i % 256) + 7
        j.numCh, j.captured_frames, j.latency_samples, j.original_length = ch, cap, lat * ch, src
        j.fs_in, j.fs_out, j.interp_kind = float(fs_in), float(fs_out), kind
        j.flags = f9.JOB_TAIL_SCAN | f9.JOB_PCM24
        j.tail_window, j.tail_hop, j.tail_required, j.tail_mode = 9600, 4800, 3, 0
        j.has_nf, j.nf_db, j.margin_pct = 1, -90.0, 0.0
        j.src_pcm, j.src_fmt, j.src_ch = pay.data_ptr() + i * ch * cap * 3, 3, ch
        j.out_pcm24 = out24.data_ptr() + i * ch * n_out * 3
    return J


variants = []
for spec in os.environ.get("VARIANTS", "sinc:0:;lagrange:1:").split(";"):
    name, kind, opts = spec.split(":")
    ctx = f9.Context(0)
    for kv in opts.split(","):
        if kv:
            ctx.set_option(kv.split("=")[0], int(kv.split("=")[1]) if "=" in kv else 1)
    variants.append((name, ctx, jobs(int(kind)), []))
R = (f9.Result * files)()
for name, ctx, J, ts in variants:
    for _ in range(3):
        assert L.f9_process_batch(ctx.handle, J, files, R) == 0
for r in range(int(os.environ.get("ROUNDS", "40"))):
    for name, ctx, J, ts in variants:
        t = time.perf_counter()
        assert L.f9_process_batch(ctx.handle, J, files, R) == 0
        ts.append(1e3 * (time.perf_counter() - t))
for name, ctx, J, ts in variants:
    print(f"{name:28s} best {min(ts):6.1f} median {np.median(ts):6.1f} mean {np.mean(ts):6.1f} max {max(ts):6.1f}  steps > 40 ms: {sum(t > 40 for t in ts):2d} of {len(ts)}   " + " ".join(f"{t:.0f}" for t in ts))
