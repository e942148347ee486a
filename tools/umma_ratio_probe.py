#!/usr/bin/env python
"""Development probe: one WindowedSinc / Lagrange plan at a given ratio and size, timed; options from the command line."""
import ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
f9 = g._load_pkg(); L = f9.lib()
fs_in, fs_out, nch, secs, kind = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4]), int(sys.argv[5])
opts = dict(a.split("=") for a in sys.argv[6:])
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(dev); torch.cuda.set_stream(stream)
n_in = int(secs * fs_in)
x = torch.randn((nch, n_in + 64), device=dev, dtype=torch.float32) * 0.25
no = f9.resampled_length(n_in, fs_in, fs_out)
y = torch.empty((nch, no), dtype=torch.float32, device=dev)
segs = (f9.ResampleSeg * nch)(*[f9.ResampleSeg(x[c].data_ptr(), 0, n_in, y[c].data_ptr(), 0, no) for c in range(nch)])
ctx = f9.Context(0); ctx.set_stream(stream.cuda_stream)
for k, v in opts.items(): ctx.set_option(k, int(v))
plan = C.c_void_p(None)
ctx._check(L.f9_resample_plan_create(ctx.handle, kind, fs_in / fs_out, segs, nch, C.byref(plan)))
ts = []
for i in range(6):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream); ctx._check(L.f9_resample_plan_run(plan)); e1.record(stream); e1.synchronize(); ts.append(e0.elapsed_time(e1))
ms = float(np.median(ts[2:]))
print(fs_in, fs_out, nch, secs, kind, opts, f"{ms:.4f} ms  frac {(4 + 4 * fs_in / fs_out) * nch * no / (ms * 1e-3) / 1e9 / 6551.4:.3f}", flush=True)
