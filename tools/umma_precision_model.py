#!/usr/bin/env python
"""CPU model of the tensor-core FIR's rounding (no GPU needed): which accumulation scheme meets 2^-20 absolute at 0 dBFS.

Model of one tcgen05.mma (kind::f16, fp32 accumulate) as measured in round 1: the 16 products of a K step are
summed exactly with the accumulator and the result is truncated toward zero to fp32.  The script evaluates, for one ratio p/q
and a test signal, the max-abs error against the sequential fp32 oracle (restated here in numpy float32, same operation order
as oracle/f9_oracle.cpp SincTraits::value) of
  * split   : round 1's scheme, x0*w0 before / after the window centre in two accumulators;
  * lobe    : x0*w0 of the taps within +-LOBE of the centre tap in their own accumulator, every other tap in another;
  * exact   : the 3-product sum without any truncation (what the fp16 split alone costs).
usage: umma_precision_model.py [p q] [amp] [periods]
"""
import sys
import numpy as np

p, q = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (320, 147)
amp = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
NP = int(sys.argv[4]) if len(sys.argv) > 4 else 1500
TAPS, NB = 200, 32
LOBE = 4

def sinc_table():
    i = np.arange(10001, dtype=np.float64)
    x = i / 100.0
    with np.errstate(divide="ignore", invalid="ignore"):
        s = np.sin(np.pi * x) / (np.pi * x)
    w = 0.5 * (1.0 + np.cos(np.pi * x / 100.0))
    t = (s * w).astype(np.float32)
    t[0] = 1.0
    t[100::100] = 0.0
    return np.concatenate([t, np.zeros(2, np.float32)])

T = sinc_table()

def tap_weights(offset):
    """float32 weights of the 200 taps at sub-sample offset (float32), as sinc_weights() in f9_tables.cpp"""
    f = np.float32
    w = np.zeros(TAPS, np.float32)
    first_frac = f(0); last = f(-1); index = 0; sign = -1
    for i in range(-100, 100):
        sp = f(f(1.0) - offset) + f(i)
        if i == -100 or (sp >= 0 and last < 0):
            idxf = f(abs(sp)) * f(100.0)
            fl = np.floor(idxf)
            index = int(fl); first_frac = f(idxf - fl); sign = -1 if sp < 0 else 1
        if sp == 0:
            v = f(1.0)
        elif -100 < sp < 100:
            v1, v2 = T[index], T[index + 1]
            v = f(v1 + f(first_frac * f(v2 - v1)))
        else:
            v = f(0)
        w[i + 100] = v
        last = sp
        index += 100 * sign
    return w

def f16(v):
    return v.astype(np.float16).astype(np.float64)

def trunc32(v):
    """float64 -> nearest-toward-zero float32, returned as float64"""
    r = v.astype(np.float32)
    over = np.abs(r.astype(np.float64)) > np.abs(v)
    r[over] = np.nextafter(r[over], np.float32(0))
    return r.astype(np.float64)

def floor16(x):
    return (x // 16) * 16

rng = np.random.default_rng(1)
n_in = NP * p + 1024
sig = sys.argv[5] if len(sys.argv) > 5 else "noise"
if sig == "noise":
    x = (rng.uniform(-1, 1, n_in) * amp).astype(np.float32)
elif sig == "sine":
    x = (amp * np.sin(2 * np.pi * 997.0 / 96000.0 * np.arange(n_in))).astype(np.float32)
else:   # square-ish burst
    x = (amp * 0.999 * np.sign(np.sin(2 * np.pi * 441.0 / 96000.0 * np.arange(n_in)))).astype(np.float32)
xs = x.astype(np.float64) * 128.0
x0 = f16(xs); x1 = f16((xs - x0) * 2048.0)

B = [(k * p) // q for k in range(q)]
G = (q + NB - 1) // NB
split_g = []
geo = []
for g in range(G):
    k0, k1 = NB * g, min(q, NB * g + NB) - 1
    wmin, wend = B[k0] - (TAPS - 1), B[k1] + 1
    t0 = floor16(wmin); ks = (wend - t0 + 15) // 16
    geo.append((t0, ks))
    centre = B[k1] - (TAPS - 1) - t0 + TAPS // 2 + 2
    split_g.append((centre + 8 + 15) // 16)
split = max(split_g)

A0 = 2                                        # first period evaluated (window stays inside x)
a = np.arange(A0, NP)
res = {k: [] for k in ("oracle", "true", "split", "lobe", "lobe3", "exact")}
for k in range(q):
    g = k // NB
    t0, ks = geo[g]
    w = tap_weights(np.float32(((k * p) % q) / q))
    w64 = w.astype(np.float64)
    w0 = f16(w64); w1 = f16((w64 - w0) * 2048.0)
    base = a * p + B[k] - (TAPS - 1)          # input index of tap 0
    idx = base[:, None] + np.arange(TAPS)[None, :]
    X = x[idx]                                # float32 [periods, taps]
    # oracle: sequential float32
    acc = np.zeros(len(a), np.float32)
    for t in range(TAPS):
        if w[t] != 0 or True:
            acc = (acc + (X[:, t] * w[t]).astype(np.float32)).astype(np.float32)
    res["oracle"].append(acc.astype(np.float64))
    res["true"].append((X.astype(np.float64) * w64[None, :]).sum(1))
    X0 = x0[idx]; X1 = x1[idx]
    sh = B[k] - (TAPS - 1) - t0               # K offset of tap 0 in the group window
    step_of_tap = (sh + np.arange(TAPS)) // 16
    centre_tap = int(np.argmax(np.abs(w)))
    is_lobe = np.abs(np.arange(TAPS) - centre_tap) <= LOBE - (1 if False else 0)
    dA = np.zeros(len(a)); dB = np.zeros(len(a)); d1 = np.zeros(len(a))
    lA = np.zeros(len(a)); lM = np.zeros(len(a)); l1 = np.zeros(len(a))
    e0 = np.zeros(len(a)); e1 = np.zeros(len(a))
    for j in range(ks):
        m = step_of_tap == j
        if not m.any():
            continue
        P0 = (X0[:, m] * w0[None, m]).sum(1)
        P1a = (X0[:, m] * w1[None, m]).sum(1)
        P1b = (X1[:, m] * w0[None, m]).sum(1)
        if j < split: dA = trunc32(dA + P0)
        else: dB = trunc32(dB + P0)
        d1 = trunc32(trunc32(d1 + P1a) + P1b)
        ml = m & is_lobe; mt = m & ~is_lobe
        if mt.any(): lA = trunc32(lA + (X0[:, mt] * w0[None, mt]).sum(1))
        if ml.any(): lM = trunc32(lM + (X0[:, ml] * w0[None, ml]).sum(1))
        l1 = d1
        e0 += P0; e1 += P1a + P1b
    f = np.float32
    def combine(a0, b0, dd1):
        s = (a0.astype(f) + b0.astype(f)).astype(f) * f(1.0 / 128.0)
        return (dd1.astype(f) * f(1.0 / (2048.0 * 128.0)) + s.astype(np.float64)).astype(f).astype(np.float64)   # fma: one rounding
    res["split"].append(combine(dA, dB, d1))
    res["lobe"].append(combine(lA, lM, l1))
    res["exact"].append((e0 + e1 / 2048.0) / 128.0)

R = {k: np.stack(v, 1) for k, v in res.items() if v}
u = 2.0 ** -20
print(f"ratio {p}/{q} amp {amp} signal {sig}: {R['oracle'].size} outputs, split step {split}, max |y| {np.abs(R['true']).max():.3f}")
for k in ("split", "lobe", "exact"):
    print(f"  {k:6s} vs oracle: max {np.abs(R[k] - R['oracle']).max() / u:.3f} x 2^-20   vs true: max {np.abs(R[k] - R['true']).max() / u:.3f}  rms {np.sqrt(np.mean((R[k] - R['true']) ** 2)) / u:.4f}")
print(f"  oracle vs true: max {np.abs(R['oracle'] - R['true']).max() / u:.3f} x 2^-20  rms {np.sqrt(np.mean((R['oracle'] - R['true']) ** 2)) / u:.4f}")
