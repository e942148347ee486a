#!/usr/bin/env python
"""Does the step-to-step variance of the e2e leg come with the NUMBER of operations?  Interleaves, in one process: the leg's bytes as
two plain copies; as 256 uploads + 256 downloads on two streams; the same with an event hand-off per group of 9 files (the pipeline's
shape without its kernels); and f9_process_batch itself."""
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
ub, db = ch * cap * 3, ch * n_out * 3
pay = torch.zeros(files * ub, dtype=torch.uint8, pin_memory=True)
out24 = torch.zeros(files * db, dtype=torch.uint8, pin_memory=True)
d_up = torch.empty(files * ub, dtype=torch.uint8, device="cuda"); d_dn = torch.empty(files * db, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
J = (f9.Job * files)()
for i in range(files):
    j = J[i]
    j.numCh, j.captured_frames, j.latency_samples, j.original_length = ch, cap, (128 * (i % 256) + 7) * ch, src
    j.fs_in, j.fs_out, j.interp_kind = float(fs_in), float(fs_out), 0
    j.flags = f9.JOB_TAIL_SCAN | f9.JOB_PCM24
    j.tail_window, j.tail_hop, j.tail_required, j.tail_mode = 9600, 4800, 3, 0
    j.has_nf, j.nf_db, j.margin_pct = 1, -90.0, 0.0
    j.src_pcm, j.src_fmt, j.src_ch = pay.data_ptr() + i * ub, 3, ch
    j.out_pcm24 = out24.data_ptr() + i * db
R = (f9.Result * files)()
ctx = f9.Context(0)
for _ in range(3):
    assert L.f9_process_batch(ctx.handle, J, files, R) == 0


def two():
    with torch.cuda.stream(s1):
        d_up.copy_(pay, non_blocking=True)
    with torch.cuda.stream(s2):
        out24.copy_(d_dn, non_blocking=True)
    s1.synchronize(); s2.synchronize()


def many(events):
    evs = []
    for k in range(0, files, 9):
        with torch.cuda.stream(s1):
            for i in range(k, min(files, k + 9)):
                d_up[i * ub:(i + 1) * ub].copy_(pay[i * ub:(i + 1) * ub], non_blocking=True)
            if events:
                e = torch.cuda.Event(); e.record(s1); evs.append(e)
        with torch.cuda.stream(s2):
            if events:
                s2.wait_event(evs[-1])
            for i in range(k, min(files, k + 9)):
                out24[i * db:(i + 1) * db].copy_(d_dn[i * db:(i + 1) * db], non_blocking=True)
    s1.synchronize(); s2.synchronize()


def lib():
    assert L.f9_process_batch(ctx.handle, J, files, R) == 0


variants = [("two copies", two, []), ("512 copies", lambda: many(False), []), ("512 copies + events", lambda: many(True), []), ("f9_process_batch", lib, [])]
for name, fn, ts in variants:
    fn()
for r in range(int(os.environ.get("ROUNDS", "40"))):
    for name, fn, ts in variants:
        torch.cuda.synchronize()
        t = time.perf_counter(); fn(); ts.append(1e3 * (time.perf_counter() - t))
for name, fn, ts in variants:
    print(f"{name:22s} best {min(ts):6.1f} median {np.median(ts):6.1f} mean {np.mean(ts):6.1f} max {max(ts):6.1f}  > 1.2 x best: {sum(t > 1.2 * min(ts) for t in ts):2d} of {len(ts)}   " + " ".join(f"{t:.0f}" for t in ts))
