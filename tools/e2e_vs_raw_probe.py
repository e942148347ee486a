#!/usr/bin/env python
"""Is the e2e leg's step-to-step variance the library's or the host's?  Alternates, in one process, a raw pinned-memory transfer of the
leg's traffic mix (1.6 GB up + 0.68 GB down as two cudaMemcpyAsync on two streams) with one f9_process_batch step over config 2's 256
file payloads, forty times, and prints both times per round."""
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
d_up = torch.empty(pay.numel(), dtype=torch.uint8, device="cuda"); d_dn = torch.empty(out24.numel(), dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
J = (f9.Job * files)()
for i in range(files):
    j = J[i]
    lat = 128 * (i % 256) + 7
    j.numCh, j.captured_frames, j.latency_samples, j.original_length = ch, cap, lat * ch, src
    j.fs_in, j.fs_out, j.interp_kind = float(fs_in), float(fs_out), 0
    j.flags = f9.JOB_TAIL_SCAN | f9.JOB_PCM24
    j.tail_window, j.tail_hop, j.tail_required, j.tail_mode = 9600, 4800, 3, 0
    j.has_nf, j.nf_db, j.margin_pct = 1, -90.0, 0.0
    j.src_pcm, j.src_fmt, j.src_ch = pay.data_ptr() + i * ch * cap * 3, 3, ch
    j.out_pcm24 = out24.data_ptr() + i * ch * n_out * 3
R = (f9.Result * files)()
ctx = f9.Context(0)
for _ in range(3):
    assert L.f9_process_batch(ctx.handle, J, files, R) == 0
raw, e2e = [], []
for r in range(40):
    torch.cuda.synchronize()
    t = time.perf_counter()
    with torch.cuda.stream(s1):
        d_up.copy_(pay, non_blocking=True)
    with torch.cuda.stream(s2):
        out24.copy_(d_dn, non_blocking=True)
    s1.synchronize(); s2.synchronize()
    raw.append(1e3 * (time.perf_counter() - t))
    t = time.perf_counter()
    assert L.f9_process_batch(ctx.handle, J, files, R) == 0
    e2e.append(1e3 * (time.perf_counter() - t))
print("raw two-copy transfer ms:", " ".join(f"{v:.1f}" for v in raw))
print("f9_process_batch     ms:", " ".join(f"{v:.1f}" for v in e2e))
print(f"raw: best {min(raw):.1f} median {np.median(raw):.1f} mean {np.mean(raw):.1f} max {max(raw):.1f}   e2e: best {min(e2e):.1f} median {np.median(e2e):.1f} mean {np.mean(e2e):.1f} max {max(e2e):.1f}")
print(f"correlation of the two series: {np.corrcoef(raw, e2e)[0, 1]:.2f}")
