#!/usr/bin/env python
"""Timeline of umma_fir_kernel's roles from the event trace of a -DF9_DIAG build.

    make -C f9-juce-resampler-studio_b200 DIAG=1
    F9DSP_DIAG_LIB=1 F9_UMMA_TRACE=4 F9_UMMA_TRACE_FILE=gpurun_out/umma_trace.bin python tools/rate_bench.py 0 96000:44100   # on the GPU
    python tools/umma_trace.py gpurun_out/umma_trace.bin                                                                 # anywhere

Events (f9_umma.cu, TR_EV): producer 1 = ring position free, box issued; converter 10 = box landed, 11 = converted (ring position
released), 12 = operand slot free, 13 = operand stored (arrive); issuer 20 = operand ready seen, 21 = stage's MMAs issued + commit,
22/23 = wait for an accumulator (before / after); epilogue 30 = group done seen, 31 = group stored."""
import sys
import numpy as np

CAP, WARPS = 1024, 18
ISS, CONV = range(5, 10), range(10, 18)        # warps 0-3 epilogue, 4 producer, 5-9 issuers, 10-17 converters (two teams of four)
raw = np.fromfile(sys.argv[1], dtype=np.int64).reshape(2, WARPS, CAP)
cta = int(sys.argv[2]) if len(sys.argv) > 2 else 0
sync = [int(raw[c, 4, CAP - 1] & 0xffffffffff) for c in range(2)]       # each CTA's clock as it left the start-up cluster barrier
raw[:, 4, CAP - 1] = 0


def load(c):
    out = {}
    for w in range(WARPS):
        r = raw[c, w]
        r = r[r != 0]
        out[w] = [(int(x >> 56) & 0xff, int(x >> 40) & 0xffff, int(x & 0xffffffffff) - sync[c] + sync[0]) for x in r
                  if (int(x >> 56) & 0xff) in (1, 10, 11, 12, 13, 14, 20, 21, 22, 23, 30, 31)]
    return out


ev = load(cta)
t0 = min(e[2] for w in ev for e in ev[w]) if any(ev[w] for w in ev) else 0


def rel(w):
    return [(e, i, t - t0) for e, i, t in ev[w]]


def deltas(w, a, b):
    """Times from event a to the next event b of warp w."""
    out, last = [], None
    for e, i, t in ev[w]:
        if e == a:
            last = t
        elif e == b and last is not None:
            out.append(t - last); last = None
    return np.array(out) if out else np.array([0])


def period(w, a):
    ts = np.array([t for e, i, t in ev[w] if e == a])
    return np.diff(ts) if len(ts) > 1 else np.array([0])


def s(x):
    return f"mean {x.mean():7.0f}  med {np.median(x):7.0f}  p90 {np.percentile(x, 90):7.0f}  max {x.max():7.0f}  n {len(x)}"


print(f"CTA {cta}: clocks relative to the first event; all figures in SM clocks")
print("producer   box-to-box              ", s(period(4, 1)))
for w in CONV:
    if not ev[w]:
        continue
    print(f"converter {w}: stage period        ", s(period(w, 10)))
    print(f"             arrive (13 -> 14)    ", s(deltas(w, 13, 14)))
    print(f"             box wait (14 -> 10)  ", s(deltas(w, 14, 10)))
    print(f"             convert (10 -> 11)   ", s(deltas(w, 10, 11)))
    print(f"             slot wait (11 -> 12) ", s(deltas(w, 11, 12)))
    print(f"             store (12 -> 13)     ", s(deltas(w, 12, 13)))
for w in ISS:
    if not ev[w]:
        continue
    print(f"issuer {w}:   stage period         ", s(period(w, 20)))
    print(f"             ready wait (21 -> 20)", s(deltas(w, 21, 20)))
    print(f"             issue (20 -> 21)     ", s(deltas(w, 20, 21)))
    print(f"             acc wait (22 -> 23)  ", s(deltas(w, 22, 23)))
for w in range(0, 4):
    if not ev[w]:
        continue
    print(f"epilogue {w}: group wait (31 -> 30)", s(deltas(w, 31, 30)))
    print(f"             drain+store (30->31) ", s(deltas(w, 30, 31)))
if len(sys.argv) > 3:          # full merged timeline
    allv = sorted((t - t0, w, e, i) for w in ev for e, i, t in ev[w])
    for t, w, e, i in allv:
        print(f"{t:9d}  warp {w:2d}  ev {e:2d}  idx {i}")

# ---- dependency gaps (same CTA): which hand-off does each wait end on? ---------------------------------------------
def first(evno, tile_stage):
    """(tile-local) stage -> {warp: [times]} for an event; stages repeat per tile, so keep occurrence order."""
    out = {}
    for w in ev:
        for e, i, t in ev[w]:
            if e == evno:
                out.setdefault(w, []).append((i, t))
    return out


def seq(evno, warps):
    """Global stage order per warp: list of times in occurrence order."""
    return {w: [t for e, i, t in ev[w] if e == evno] for w in warps if ev[w]}


conv = [w for w in range(10, 18) if ev[w]]
iss = [w for w in range(5, 10) if ev[w]]
if conv and iss:
    st_done = seq(13, conv); slot_seen = seq(12, conv); full_seen = seq(10, conv); conv_done = seq(11, conv)
    ready = seq(20, iss); issued = seq(21, iss)
    n_st = min(len(v) for v in ready.values())
    # stage g (global, from the first traced stage) is converted by team g % 2: warps 10-13 (team 0) or 14-17 (team 1)
    team = {0: [w for w in conv if w < 14], 1: [w for w in conv if w >= 14]}
    first_team = 0 if min(st_done[w][0] for w in team[0]) < min(st_done[w][0] for w in team[1]) else 1
    gaps_ready, gaps_slot, spans_issue, spans_st = [], [], [], []
    for g in range(n_st):
        tm = team[(g + first_team) % 2]
        k = g // 2
        if any(k >= len(st_done[w]) for w in tm):
            break
        last_st = max(st_done[w][k] for w in tm)                      # this CTA's last converter arrive of stage g
        first_ready = min(ready[w][g] for w in iss)
        gaps_ready.append(first_ready - last_st)
        last_issue = max(issued[w][g] for w in iss)
        spans_issue.append(last_issue - first_ready)
        if g + 4 < n_st:
            tm4 = team[(g + 4 + first_team) % 2]
            k4 = (g + 4) // 2
            if all(k4 < len(slot_seen[w]) for w in tm4):
                gaps_slot.append(min(slot_seen[w][k4] for w in tm4) - last_issue)     # slot reuse: stage g + 4 after stage g's commit
    print("own CTA's last operand store of a stage -> first issuer sees 'ready'  ", s(np.array(gaps_ready)), "(negative: the peer CTA was later)")
    print("first 'ready' seen -> last issuer done with the stage (issue + commit) ", s(np.array(spans_issue)))
    print("last issuer commit of stage g -> first converter sees slot free (g + 4)", s(np.array(gaps_slot)))

# ---- across the pair (clocks aligned at the start-up cluster barrier) ---------------------------------------------------
if sync[0] and sync[1]:
    peer = load(1 - cta)
    lead, follow = (ev, peer) if cta == 0 else (peer, ev)
    cw = [w for w in range(10, 18) if lead[w] and follow[w]]
    iw = [w for w in range(5, 10) if lead[w]]
    if cw and iw:
        def sq(d, evno, warps):
            return {w: [t for e, i, t in d[w] if e == evno] for w in warps}
        l_st, f_st = sq(lead, 13, cw), sq(follow, 13, cw)
        l_slot, f_slot = sq(lead, 12, cw), sq(follow, 12, cw)
        rdy, isd = sq(lead, 20, iw), sq(lead, 21, iw)
        n = min(len(v) for v in rdy.values())
        tm = {0: [w for w in cw if w < 14], 1: [w for w in cw if w >= 14]}
        ft = 0 if min(l_st[w][0] for w in tm[0]) < min(l_st[w][0] for w in tm[1]) else 1
        a, b, c, d, e2 = [], [], [], [], []
        for g in range(n):
            t = tm[(g + ft) % 2]; k = g // 2
            if any(k >= len(l_st[w]) or k >= len(f_st[w]) for w in t):
                break
            ls, fs = max(l_st[w][k] for w in t), max(f_st[w][k] for w in t)
            r0 = min(rdy[w][g] for w in iw)
            a.append(fs - ls); b.append(r0 - max(ls, fs))
            li = max(isd[w][g] for w in iw)
            if g + 4 < n:
                t4 = tm[(g + 4 + ft) % 2]; k4 = (g + 4) // 2
                if all(k4 < len(l_slot[w]) and k4 < len(f_slot[w]) for w in t4):
                    c.append(min(l_slot[w][k4] for w in t4) - li); d.append(min(f_slot[w][k4] for w in t4) - li)
                    e2.append(max(max(l_st[w][k4] for w in t4), max(f_st[w][k4] for w in t4)) - max(ls, fs))
        print("follower's last operand store - leader's (same stage)                    ", s(np.array(a)))
        print("later of the two stores -> first issuer sees 'ready'                      ", s(np.array(b)))
        print("last issuer commit of stage g -> LEADER's converters see the slot free    ", s(np.array(c)))
        print("last issuer commit of stage g -> FOLLOWER's converters see the slot free  ", s(np.array(d)))
        print("operand store of stage g -> operand store of stage g + 4 (the ring's loop)", s(np.array(e2)))
