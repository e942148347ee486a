#!/usr/bin/env python
"""Development check of the tensor-core FIR against the oracle: max-abs error and SNR per ratio / kind / signal."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
f9 = g._load_pkg()
from oracle import oracle as O

def snr_db(ref, got):
    err = np.sqrt(np.mean((ref.astype(np.float64) - got.astype(np.float64)) ** 2))
    sig = np.sqrt(np.mean(ref.astype(np.float64) ** 2))
    return np.inf if err == 0 else 20 * np.log10(sig / err)

ctx = f9.Context(0)
rng = np.random.default_rng(0)
cases = [(96000, 44100), (44100, 48000), (48000, 192000), (48000, 96000), (24000, 192000), (12000, 192000), (96000, 48000), (192000, 48000), (88200, 48000), (44100, 96000), (48000, 44100)]
n_in = int(sys.argv[1]) if len(sys.argv) > 1 else 60000
bad = 0
for kind in (0, 1, 2, 3):
    for fs_in, fs_out in cases:
        for amp in (0.5, 1e-3):
            x = (rng.uniform(-1, 1, (2, n_in)) * amp).astype(np.float32)
            t0 = time.time()
            y = ctx.resample(x, fs_in, fs_out, kind)
            n_out = y.shape[1]
            ref = np.stack([O.resample_channel(kind, fs_in / fs_out, x[c], n_out)[0] for c in range(2)])
            err = float(np.max(np.abs(y - ref))); s = snr_db(ref, y)
            ok = err <= 2.0 ** -20 * max(amp * 2, 1e-3) * 2 and s >= 120
            bad += not ok
            print(f"kind {kind} {fs_in}->{fs_out} amp {amp:g}: max err {err:.3e} ({err / 2.0 ** -20:.3f} x 2^-20)  snr {s:.1f} dB {'ok' if ok else 'BAD'}", flush=True)
# out-of-range input -> fp32 redo
x = (rng.uniform(-1, 1, (1, 20000))).astype(np.float32); x[0, 5000] = 70000.0
y = ctx.resample(x, 96000, 44100, 0)
ref = O.resample_channel(0, 96000 / 44100, x[0], y.shape[1])[0]
print("overflow redo: max err", float(np.max(np.abs(y[0] - ref))))
print("launches", ctx.launch_count, "bad", bad)
sys.exit(1 if bad else 0)
