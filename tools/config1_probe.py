#!/usr/bin/env python
"""Config 1 (one 60 s stereo file, 44.1 -> 48 kHz) through a plan: where its ~40 us go.  Run plain for CUDA-event times of
f9_resample_plan_run (L2 flushed between repetitions and not), or under `ncu --metrics gpu__time_duration.sum` for the launches."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

f9 = g._load_pkg()
if os.environ.get("F9DSP_DIAG_LIB"):          # development: the -DF9_DIAG build (make -C f9-juce-resampler-studio_b200 DIAG=1)
    f9.LIB_PATH = os.path.join(os.path.dirname(os.path.dirname(f9.LIB_PATH)), "lib_diag", "libf9dsp.so")
dev = torch.device("cuda", 0)
ctx = f9.Context(0)
stream = torch.cuda.Stream(dev); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
L = f9.lib()
gen = torch.Generator(device=dev); gen.manual_seed(1)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
nch, n_in, fs_in, fs_out = 2, 60 * 44100, 44100, 48000
x = torch.randn((nch, n_in), generator=gen, device=dev, dtype=torch.float32) * 0.25
no = f9.resampled_length(n_in, fs_in, fs_out)
y = torch.empty((nch, no), dtype=torch.float32, device=dev)
segs = (f9.ResampleSeg * nch)(*[f9.ResampleSeg(x[c].data_ptr(), 0, n_in, y[c].data_ptr(), 0, no) for c in range(nch)])
for kind in (0, 1):
    plan = C.c_void_p(None)
    ctx._check(L.f9_resample_plan_create(ctx.handle, kind, fs_in / fs_out, segs, nch, C.byref(plan)))
    for do_flush in (True, False):
        for _ in range(3):
            ctx._check(L.f9_resample_plan_run(plan))
        ts = []
        for _ in range(9):
            if do_flush:
                flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream); ctx._check(L.f9_resample_plan_run(plan)); e1.record(stream); e1.synchronize(); ts.append(e0.elapsed_time(e1))
        ts.sort()
        print(f"kind {kind} {'L2 flushed' if do_flush else 'L2 warm   '}: {ts[len(ts) // 2] * 1e3:.1f} us per f9_resample_plan_run  (roofline {4.0 * nch * (n_in + no) / 6551.4e9 * 1e6:.1f} us)", flush=True)
    L.f9_plan_destroy(plan)
