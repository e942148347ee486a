#!/usr/bin/env python
"""Stall-reason totals and hottest SASS lines from an ncu report's source page: ncu_stalls.py <ncu-rep> [top] [section]"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
sec = int(sys.argv[3]) if len(sys.argv) > 3 else 0
lo = starts[sec]; hi = starts[sec + 1] if sec + 1 < len(starts) else len(rows)
print(rows[lo][1][:120])
hdr = rows[lo + 1]; ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = {s: 0 for s in stalls}; data = []
for r in rows[lo + 2:hi]:
    if len(r) < len(hdr): continue
    data.append((int(r[ix["# Samples"]] or 0), r))
    for s in stalls: tot[s] += int(r[ix[s]] or 0)
T = sum(tot.values()) or 1
print("total samples", T, " instructions executed", sum(int(r[ix["Instructions Executed"]] or 0) for _, r in data))
for s in sorted(stalls, key=lambda s: -tot[s])[:8]: print(f"  {s:26s} {tot[s]:9d} {100 * tot[s] / T:5.1f}%")
top = int(sys.argv[2]) if len(sys.argv) > 2 else 20
for samp, r in sorted(data, key=lambda x: -x[0])[:top]:
    t2 = sorted(stalls, key=lambda s: -int(r[ix[s]] or 0))[:2]
    print(f"{samp:8d} {r[ix['Address']][-5:]} {r[ix['Source']].strip()[:64]:64s} exec={r[ix['Instructions Executed']]:>9s} {t2[0][6:]}={r[ix[t2[0]]]} {t2[1][6:]}={r[ix[t2[1]]]}")
