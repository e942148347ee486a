#!/usr/bin/env python
"""Development probe of the file-bytes e2e leg (config 2 shape): f9_process_batch over 256 stereo 24-bit payloads in pinned memory,
swept over the chunk size of the two-slot pipeline (option F9_BATCH_CHUNK_MB) and input / output forms."""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if os.environ.get("BIND"):          # run (and first-touch the pinned buffers) on the CPUs of GPU 0's NUMA node
    import pynvml
    pynvml.nvmlInit()
    words = (os.cpu_count() + 63) // 64
    mask = pynvml.nvmlDeviceGetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(0), words)
    cpus = {64 * w + b for w in range(words) for b in range(64) if (mask[w] >> b) & 1} & os.sched_getaffinity(0)
    print("bind: gpu0 cpus", sorted(cpus), "of", sorted(os.sched_getaffinity(0)), flush=True)
    if cpus:
        os.sched_setaffinity(0, cpus)
import __graft_entry__ as g
f9 = g._load_pkg()
L = f9.lib()

files, ch, fs_in, fs_out, src = int(os.environ.get("FILES", "256")), 2, 96000, 44100, 960000
cap = (src + 128 * 255 + 7 + 48000 + 63) // 64 * 64
n_out = f9.resampled_length(src, fs_in, fs_out)
pay = torch.zeros(files * ch * cap * 3, dtype=torch.uint8, pin_memory=True)
out24 = torch.zeros(files * ch * n_out * 3, dtype=torch.uint8, pin_memory=True)
capf = torch.zeros((files, ch, cap), dtype=torch.float32, pin_memory=True)
outf = torch.zeros((files, ch, n_out), dtype=torch.float32, pin_memory=True)
fp = C.POINTER(C.c_float)


def jobs_for(mode):
    J = (f9.Job * files)()
    keep = []
    for i in range(files):
        j = J[i]
        lat = 128 * (i % 256) + 7
        j.numCh, j.captured_frames, j.latency_samples, j.original_length = ch, cap, lat * ch, src
        j.fs_in, j.fs_out, j.interp_kind = float(fs_in), float(fs_out), int(os.environ.get("KIND", "0"))
        j.flags = f9.JOB_TAIL_SCAN if not os.environ.get("NOTAIL") else 0
        j.tail_window, j.tail_hop, j.tail_required, j.tail_mode = 9600, 4800, 3, 0
        j.has_nf, j.nf_db, j.margin_pct = 1, -90.0, 0.0
        if mode[0] == "p":
            j.src_pcm, j.src_fmt, j.src_ch = pay.data_ptr() + i * ch * cap * 3, 3, ch
        else:
            a = (fp * ch)(*[C.cast(capf.data_ptr() + 4 * (i * ch + c) * cap, fp) for c in range(ch)]); keep.append(a); j.captured = a
        if mode[1] == "p":
            j.flags |= f9.JOB_PCM24; j.out_pcm24 = out24.data_ptr() + i * ch * n_out * 3
        else:
            a = (fp * ch)(*[C.cast(outf.data_ptr() + 4 * (i * ch + c) * n_out, fp) for c in range(ch)]); keep.append(a); j.out = a; j.out_capacity = n_out
    return J, keep


for mode in (os.environ.get("MODES", "pp,ff,pf,fp").split(",")):
    J, keep = jobs_for(mode)
    R = (f9.Result * files)()
    for chunk in [int(c) for c in os.environ.get("CHUNKS", "16,32,64,128,256,512").split(",")]:
        ctx = f9.Context(0)
        ctx.set_option("F9_BATCH_CHUNK_MB", chunk)
        for kv in os.environ.get("OPTS", "").split(","):
            if kv:
                ctx.set_option(kv.split("=")[0], int(kv.split("=")[1]) if "=" in kv else 1)
        for _ in range(2):
            assert L.f9_process_batch(ctx.handle, J, files, R) == 0
        ts = []
        for _ in range(int(os.environ.get("REPS", "4"))):
            t = time.perf_counter(); assert L.f9_process_batch(ctx.handle, J, files, R) == 0; ts.append(1e3 * (time.perf_counter() - t))
        print(f"in/out {mode} chunk {chunk:4d} MB: best {min(ts):7.2f} ms  mean {np.mean(ts):7.2f} ms   ({files * ch * n_out / min(ts) / 1e6:.2f} Gsamples/s)  each "
              + " ".join(f"{t:.1f}" for t in ts), flush=True)
        ctx.close()
