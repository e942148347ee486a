#!/usr/bin/env python
"""Cost of calculateRMS's reference-order re-sum (option F9_RMS_FORCE_ORDER) on ONE capture, against the tree sum of the same
capture: what a flagged buffer adds to a batch (f9_scan.cu, sum_squares_in_reference_order)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

f9 = g._load_pkg()
dev = torch.device("cuda", 0)
ctx = f9.Context(0)
stream = torch.cuda.Stream(dev); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
L = f9.lib()
gen = torch.Generator(device=dev); gen.manual_seed(1)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for ch, frames, amp, label in ((2, 240000, 1e-4, "5 s stereo, noise 1e-4 + impulse"), (2, 240000, 0.25, "5 s stereo, noise 0.25"), (2, 960000, 0.25, "10 s stereo 96 k, noise 0.25")):
    rec = torch.randn((1, ch, frames), generator=gen, device=dev, dtype=torch.float32) * amp
    if amp < 1e-3:
        rec[0, 0, 4321] = 0.9
    bufs = (f9.DevBuffer * 1)(f9.DevBuffer(rec[0].data_ptr(), frames, ch, frames))
    sumsq = torch.empty(1, dtype=torch.float64, device=dev); pk = torch.empty(1, dtype=torch.float32, device=dev)
    for opt in ("F9_RMS_TREE_SUM", "F9_RMS_FORCE_ORDER"):
        ctx.clear_options(); ctx.set_option(opt, 1)
        ts = []
        for _ in range(8):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream); ctx._check(L.f9_dev_stats_batch(ctx.handle, bufs, 1, sumsq.data_ptr(), pk.data_ptr())); e1.record(stream); e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        print(f"{label}: {opt}: {ts[len(ts) // 2] * 1e3:.1f} us  (sum of squares {sumsq.item():.17g})", flush=True)
