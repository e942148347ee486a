#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on BASELINE.json's configs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME] [--files F] [--scaling weak|strong]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...          (N > 1, one rank per GPU)

metric  : resampled output Msamples/s, all channels (BASELINE.json "metric")
workload: configs[1] by default -- batch of 256 stereo files 96 kHz -> 44.1 kHz with latency trim and tail-silence detection,
          10 s per file, WindowedSinc (the Lagrange numbers ride along in "lagrange").  --workload selects config 1, config 3
          (64 channels, 10 min, 48 -> 192 k, time-segmented with halos) or config 5 (4096 mixed-rate files to 48 k).
step    : one pass of the hot path over the rank's share of the workload: reverb-tail scan of every capture + trimLatency fused
          into the polyphase resampler, captures resident in HBM (value); or file bytes in pinned host memory through the batch
          job flow -- 24-bit payload up, 24-bit payload down, deinterleave / int<->float on the device (e2e).
N > 1   : config 2 is weak scaling by default (every rank owns its own 256 files); --scaling strong, and configs 3 / 5 always, shard
          ONE workload over the ranks with the product's partition (f9_multi_partition: files, channel groups, time segments
          with halos; greedy packing by output samples).  No data-path collective; time = max over ranks.
--impl reference : the reference's CPU implementation of the path.  JUCE is not vendored by the reference and the reference
          cannot be compiled here, so this times the oracle port (oracle/, scalar C++, one interpolator object per channel as JUCE
          runs it) on all host cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import ctypes as C
import importlib.util
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "f9-juce-resampler-studio_b200")
sys.path.insert(0, ROOT)


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


W = _load("f9workloads", os.path.join(PKG, "py", "workloads.py"))


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "sm_max_mhz": 1965.0}, "fallback"


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.idx = device_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0: float, t1: float):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        rows = [l for (ts, l) in self.lines if t0 - 0.05 <= ts <= t1 + 0.2] or [l for (_, l) in self.lines]
        for l in rows:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU baseline (oracle port)
_CPU_CAPS = {}


def cpu_inputs(workload: str, files: int):
    """The CPU sample's captures, generated once per (workload, file count) and reused by every step (generation is not timed)."""
    key = (workload, files)
    if key not in _CPU_CAPS:
        secs = 20.0 if workload.startswith("config3") else None              # config 3: 20 s of the 10 minutes as the sample
        specs = W.file_specs(workload, files, 0, seconds=secs)
        _CPU_CAPS[key] = (specs, [W.window(s, 0, s.num_ch, 0, s.cap_frames) for s in specs])
    return _CPU_CAPS[key]


def cpu_leg(workload: str, kind: int, files: int, threads: int):
    """One bounded sample of the workload on the host cores: tail scan + trimLatency + conversion per file (whatever the config
    asks for), files statically partitioned over `threads` (ctypes calls release the GIL).  Returns (out_samples, seconds)."""
    from oracle import oracle as O

    specs, caps = cpu_inputs(workload, files)
    O.lib()
    total = [0]

    def one(i):
        s = specs[i]
        if s.tail:
            O.tail_scan(caps[i], s.src_frames + s.latency, int(s.fs_in * 0.1), int(s.fs_in * 0.05), 3, 0, True, -90.0, 0.0)
        trimmed, _ = O.trim_latency(caps[i], s.latency * s.num_ch, s.src_frames)
        n_out = -((-s.src_frames * s.fs_out) // s.fs_in)
        if s.fs_in != s.fs_out:
            O.resample_channels(kind, s.fs_in / s.fs_out, trimmed, n_out, threads=1)
        return s.num_ch * n_out

    # work items: files, or the channels of a multichannel file
    items = []
    for i, s in enumerate(specs):
        items.append(i)
    threads = max(1, threads)
    if len(specs) == 1 and specs[0].num_ch > 1 and threads > 1:
        s = specs[0]
        n_out = -((-s.src_frames * s.fs_out) // s.fs_in)
        t0 = time.perf_counter()
        O.resample_channels(kind, s.fs_in / s.fs_out, caps[0], n_out, threads=min(threads, s.num_ch))
        return s.num_ch * n_out, time.perf_counter() - t0
    threads = min(threads, len(items))
    bounds = [len(items) * k // threads for k in range(threads + 1)]
    done = [0] * threads

    def work(k):
        for i in items[bounds[k]:bounds[k + 1]]:
            done[k] += one(i)

    ts = [threading.Thread(target=work, args=(k,)) for k in range(threads)]
    t0 = time.perf_counter()
    [t.start() for t in ts]
    [t.join() for t in ts]
    dt = time.perf_counter() - t0
    total[0] = sum(done)
    return total[0], dt


def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    nfiles = W.CONFIGS[args.workload][4]
    files = max(1, min(nfiles, max(cores, args.ref_files)))
    for _ in range(args.warmup):
        cpu_leg(args.workload, 0, files, cores)
    total, secs = 0, 0.0
    for _ in range(args.steps):
        n, dt = cpu_leg(args.workload, 0, files, cores)
        total += n; secs += dt
    v = total / secs / 1e6
    s0 = W.file_specs(args.workload, 1)[0]
    line = {"impl": "reference", "metric": "resampled output Msamples/s (all channels)", "value": v, "unit": "Msamples/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "interpolator": "WindowedSinc", "files_in_workload": nfiles,
                       "fs_in": s0.fs_in if not args.workload.startswith("config5") else "mixed", "fs_out": s0.fs_out, "channels": s0.num_ch,
                       "seconds_per_file": s0.src_frames / s0.fs_in},
            "cpu_baseline": {"value": v, "unit": "Msamples/s", "cores": cores, "kind": "port",
                             "sample": f"{files} of {nfiles} files per step (tail scan + trimLatency + WindowedSinc), oracle port, g++ -O2 -ffp-contract=off"},
            "e2e": {"value": v, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# The driver reads ONE JSON line from stdout.  Libraries write there too (NCCL prints its version banner on stdout at
# NCCL_DEBUG=VERSION and WARN alike), so file descriptor 1 is pointed at stderr for the life of the process and the line goes
# out through a private duplicate of the original stdout.
_REAL_STDOUT = None


def claim_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel, from the ncu --set full capture of this same
# command (not measured by this run: a number taken under a profiler is never a bench value, and ncu is not available inside a
# timed run).  Keyed by (workload, files per GPU); None when no capture of that shape is committed.
TRAFFIC = {
    ("config2_256x_stereo_96k_to_44k1_trim_tail", 256): (2.035183e9 + 873.595136e6, "profiles/r02d_umma_full.txt (ncu --set full of this command at 256 files, round 2's kernel: dram__bytes_read + dram__bytes_write of the FIR launch)"),
}


# ------------------------------------------------------------------------------------------------ GPU arm
def bind_to_gpu_numa_node(local_rank: int):
    """One process per GPU: run on the CPUs next to this rank's GPU, so that the pinned host buffers of the e2e leg are
    allocated (first touch) on the NUMA node its PCIe link hangs off.  Returns the CPU count bound to, or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w in range(words) for b in range(64) if (mask[w] >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def kernel_name(kind: int, fs_in: int, fs_out: int) -> str:
    from math import gcd
    g = gcd(fs_in, fs_out)
    p, q = fs_in // g, fs_out // g
    if kind == 0:
        if p == 1 and q in (2, 4, 8, 16):
            return "hankel_fir_kernel (WindowedSinc at integer upsampling on tcgen05: Hankel operand, weights resident in TMEM)"
        return "umma_fir_kernel (WindowedSinc polyphase FIR on tcgen05: TMA-fed CTA pairs, fp16 2-split, fp32 TMEM accumulators)"
    if q == 1 and p >= 2:
        return "umma_fir_kernel (Lagrange at integer decimation on tcgen05)"
    return "short_kernel (Lagrange polyphase FIR on CUDA cores, fp32: cp.async-staged tiles, slot weights in registers)"


def pcie_probe(torch, dev, barrier, world, dist, nbytes=1 << 30):
    """Pinned-memory DMA rates of this box, measured in this run on every rank AT THE SAME TIME (the host side is shared):
    H2D alone, D2H alone, both at once.  GB/s per rank (rank 0's) and summed over ranks."""
    h_up = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True); d_up = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    h_dn = torch.empty(nbytes // 2, dtype=torch.uint8, pin_memory=True); d_dn = torch.empty(nbytes // 2, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def run(up, dn):
        barrier()
        t0 = time.perf_counter()
        if up:
            with torch.cuda.stream(s1):
                d_up.copy_(h_up, non_blocking=True)
        if dn:
            with torch.cuda.stream(s2):
                h_dn.copy_(d_dn, non_blocking=True)
        s1.synchronize(); s2.synchronize()
        return time.perf_counter() - t0

    run(True, True)
    out = {}
    for name, up, dn in (("h2d", True, False), ("d2h", False, True), ("both", True, True)):
        dt = min(run(up, dn) for _ in range(2))
        gbs = ((nbytes if up else 0) + (nbytes // 2 if dn else 0)) / dt / 1e9
        tot = gbs
        if world > 1:
            t = torch.tensor([gbs], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            tot = float(t.item())
        out[name + "_gbs"] = round(gbs, 2)
        out[name + "_gbs_all_ranks"] = round(tot, 2)
    out["note"] = "cudaMemcpyAsync from/to pinned buffers, 1 GiB up / 0.5 GiB down, all ranks at once; 'both' = the e2e leg's traffic mix"
    return out


def run_gpu(args, rank: int, local_rank: int, world: int):
    import torch
    import torch.distributed as dist

    f9 = _load("f9dsp", os.path.join(PKG, "py", "f9dsp.py"))
    numa_cpus = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL prints its version banner on stdout at NCCL_DEBUG=VERSION (the image's default): keep stdout to the one JSON line
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)
    peaks, peak_src = measured_peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    ctx = f9.Context(local_rank)                      # raises when the CUDA library or the GPU is missing
    # One explicit stream carries the library's kernels AND the timing events (the legacy default stream has
    # handle 0, which f9_set_stream reads as "use the context's own stream": events there would time nothing).
    stream = torch.cuda.Stream(dev)
    assert stream.cuda_stream != 0
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    L = f9.lib()

    # ---- the rank's share of the workload: units from the product's partition -----------------------------------------------
    strong = args.scaling == "strong" or not args.workload.startswith("config2")
    nfiles_cfg = args.files or W.CONFIGS[args.workload][4]
    if strong:
        specs = W.file_specs(args.workload, nfiles_cfg, 0, seconds=args.seconds)
    else:
        specs = W.file_specs(args.workload, nfiles_cfg, rank * nfiles_cfg, seconds=args.seconds)
    n_files = len(specs)
    Jd = (f9.Job * n_files)()                          # descriptors for the partition (shapes only; pointers filled per leg)
    dummy = (C.POINTER(C.c_float) * 64)(*[C.cast(0x1000, C.POINTER(C.c_float))] * 64)
    for i, s in enumerate(specs):
        Jd[i].captured = dummy; Jd[i].out = dummy; Jd[i].out_capacity = 1 << 30
        Jd[i].numCh, Jd[i].captured_frames = s.num_ch, s.cap_frames
        Jd[i].latency_samples, Jd[i].original_length = s.latency * s.num_ch, s.src_frames
        Jd[i].fs_in, Jd[i].fs_out, Jd[i].interp_kind = float(s.fs_in), float(s.fs_out), f9.WINDOWED_SINC
        Jd[i].flags = f9.JOB_TAIL_SCAN if s.tail else 0
        Jd[i].tail_window, Jd[i].tail_hop, Jd[i].tail_required, Jd[i].tail_mode = int(s.fs_in * 0.1), int(s.fs_in * 0.05), 3, f9.TAIL_RMS
        Jd[i].has_nf, Jd[i].nf_db, Jd[i].margin_pct = 1, -90.0, 0.0
    if strong:
        seg_out = 0
        if args.workload.startswith("config3") and world > 1:
            seg_out = max(1 << 20, f9.resampled_length(specs[0].src_frames, specs[0].fs_in, specs[0].fs_out) // 8)   # "time-segmented with halos"
        U, nu = f9.partition(Jd, n_files, world, seg_out)
        units = [U[k] for k in range(nu) if U[k].device == rank and not U[k].tail_only]
        n_units_total = nu
    else:
        U, nu = f9.partition(Jd, n_files, 1, 0)
        units = [U[k] for k in range(nu)]
        n_units_total = nu * world

    # ---- resident layout: one input and one output arena; a unit's trimmed start sits on 16 bytes (the library's own upload layout)
    class RU:       # resident unit
        pass
    rus, in_floats, out_floats = [], 0, 0
    for u in units:
        s = specs[u.job]
        r = RU(); r.u, r.s = u, s
        r.convert = s.fs_in != s.fs_out
        r.ratio = s.fs_in / s.fs_out
        copied = max(0, min(s.src_frames, s.cap_frames - s.latency))           # trimLatency arithmetic (MainComponent.cpp:833-845)
        n_out_file = f9.resampled_length(s.src_frames, s.fs_in, s.fs_out) if r.convert else s.src_frames
        if u.num_out == 0:
            r.first_abs, r.in_len, r.in_offset, r.n0, r.n_out, r.in_avail, r.lat_in = 0, s.cap_frames, 0, 0, n_out_file, copied, s.latency
            r.pad = 0 if args.unaligned else (-s.latency) % 4
        else:
            f, l = f9.segment_input_range(f9.WINDOWED_SINC, r.ratio, u.n0, u.num_out)
            f = max(0, f) // 4 * 4; l = max(f, min(copied, l))
            r.first_abs, r.in_len, r.in_offset, r.n0, r.n_out, r.in_avail, r.lat_in = s.latency + f, l - f, f, u.n0, u.num_out, l - f, 0
            r.pad = 0
        r.in_stride = (r.in_len + r.pad + 8 + 63) // 64 * 64
        r.out_stride = (r.n_out + 63) // 64 * 64
        r.in_at, r.out_at = in_floats, out_floats
        in_floats += u.num_ch * r.in_stride; out_floats += u.num_ch * r.out_stride
        rus.append(r)
    d_in = torch.zeros(max(in_floats, 1), dtype=torch.float32, device=dev)
    d_out = torch.empty(max(out_floats, 1), dtype=torch.float32, device=dev)
    # Captures are what a 24-bit file holds: generated, packed to the 24-bit payload and unpacked again ON THE DEVICE by the
    # library's own format kernels, so that the resident leg and the file-bytes e2e leg convert the very same samples.
    whole_files = all(r.u.num_out == 0 and r.u.num_ch == r.s.num_ch for r in rus)
    for r in rus:
        for c in range(r.u.num_ch):
            a = r.in_at + c * r.in_stride + r.pad
            d_in[a:a + r.in_len] = W.window(r.s, r.u.ch0 + c, 1, r.first_abs, r.in_len, dev)[0]
    payload_h = None
    if whole_files and not args.no_e2e:
        nb = [r.s.num_ch * r.s.cap_frames * 3 for r in rus]
        offs = np.concatenate([[0], np.cumsum([(b + 63) // 64 * 64 for b in nb])]).astype(np.int64)
        d_pay = torch.empty(int(offs[-1]) + 64, dtype=torch.uint8, device=dev)
        bufs_in = (f9.DevBuffer * len(rus))(*[f9.DevBuffer(d_in.data_ptr() + 4 * (r.in_at + r.pad), r.in_stride, r.s.num_ch, r.s.cap_frames) for r in rus])
        ptrs = (C.c_void_p * len(rus))(*[d_pay.data_ptr() + int(offs[i]) for i in range(len(rus))])
        ctx._check(L.f9_dev_planar_to_pcm24_batch(ctx.handle, bufs_in, ptrs, len(rus)))
        for key in sorted({r.s.num_ch for r in rus}):
            idx = [i for i, r in enumerate(rus) if r.s.num_ch == key]
            b2 = (f9.DevBuffer * len(idx))(*[bufs_in[i] for i in idx]); p2 = (C.c_void_p * len(idx))(*[ptrs[i] for i in idx])
            ctx._check(L.f9_dev_pcm_to_planar_batch(ctx.handle, p2, f9.PCM_S24LE, key, b2, len(idx)))
        ctx.synchronize()
        payload_h = torch.empty(int(offs[-1]) + 64, dtype=torch.uint8, pin_memory=True)
        payload_h.copy_(d_pay); torch.cuda.synchronize(dev)
        del d_pay

    # ---- descriptors: tail scan over whole-file units, resample segments with trimLatency fused as a pointer offset ----------
    tail_rus = [r for r in rus if r.s.tail and r.u.num_out == 0]
    tbufs = (f9.DevBuffer * max(len(tail_rus), 1))()
    tpar = (f9.TailParams * max(len(tail_rus), 1))()
    max_polls = 1
    for i, r in enumerate(tail_rus):
        s = r.s
        tbufs[i] = f9.DevBuffer(d_in.data_ptr() + 4 * (r.in_at + r.pad), r.in_stride, s.num_ch, s.cap_frames)
        tpar[i] = f9.TailParams(s.src_frames + s.latency, int(s.fs_in * 0.1), int(s.fs_in * 0.05), 3, f9.TAIL_RMS, 1, -90.0, 0.0)
        max_polls = max(max_polls, (s.cap_frames - s.src_frames) // int(s.fs_in * 0.05))
    stops = torch.empty(max(len(tail_rus), 1), dtype=torch.int64, device=dev)
    flags = torch.empty(max(len(tail_rus), 1) * max_polls, dtype=torch.int32, device=dev)
    groups = {}                                        # (fs_in, fs_out) -> list of ResampleSeg
    copies = []                                        # units without conversion: trimLatency as a copy
    for r in rus:
        for c in range(r.u.num_ch):
            in_ptr = d_in.data_ptr() + 4 * (r.in_at + c * r.in_stride + r.pad + r.lat_in)
            out_ptr = d_out.data_ptr() + 4 * (r.out_at + c * r.out_stride)
            if r.convert:
                groups.setdefault((r.s.fs_in, r.s.fs_out), []).append(f9.ResampleSeg(in_ptr, r.in_offset, r.in_avail, out_ptr, r.n0, r.n_out))
        if not r.convert:
            copies.append(r)
    plans = {}
    for kind in (f9.WINDOWED_SINC, f9.LAGRANGE):
        for key, segl in groups.items():
            arr = (f9.ResampleSeg * len(segl))(*segl)
            p = C.c_void_p(None)
            ctx._check(L.f9_resample_plan_create(ctx.handle, kind, key[0] / key[1], arr, len(segl), C.byref(p)))
            plans[(kind, key)] = p
    cbufs = (f9.DevBuffer * max(len(copies), 1))(); obufs = (f9.DevBuffer * max(len(copies), 1))(); clat = (C.c_int * max(len(copies), 1))()
    for i, r in enumerate(copies):
        cbufs[i] = f9.DevBuffer(d_in.data_ptr() + 4 * (r.in_at + r.pad), r.in_stride, r.u.num_ch, r.s.cap_frames)
        obufs[i] = f9.DevBuffer(d_out.data_ptr() + 4 * r.out_at, r.out_stride, r.u.num_ch, r.n_out)
        clat[i] = r.s.latency * r.u.num_ch
    out_by_group = {key: sum(s.num_out for s in segl) for key, segl in groups.items()}
    dominant = max(out_by_group, key=out_by_group.get) if out_by_group else None
    out_samples = sum(r.u.num_ch * r.n_out for r in rus)
    alg_dom = (4.0 + 4.0 * dominant[0] / dominant[1]) * out_by_group[dominant] if dominant else 0.0       # 4 + 4*ratio per output

    def step(kind, k0=None, k1=None):
        if tail_rus:
            ctx._check(L.f9_dev_tail_scan_batch(ctx.handle, tbufs, tpar, len(tail_rus), stops.data_ptr(), flags.data_ptr(), max_polls))
        for key in groups:
            if key == dominant and k0 is not None:
                k0.record(stream)
            ctx._check(L.f9_resample_plan_run(plans[(kind, key)]))
            if key == dominant and k1 is not None:
                k1.record(stream)
        if copies:
            ctx._check(L.f9_dev_trim_batch(ctx.handle, cbufs, clat, obufs, len(copies), 0))

    def timed(kind, steps, warmup):
        for _ in range(warmup):
            step(kind)
        barrier()
        k0 = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        k1 = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches0 = ctx.launch_count
        e0.record(stream)
        for s in range(steps):
            step(kind, k0[s], k1[s])
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        kms = [a.elapsed_time(b) for a, b in zip(k0, k1)] if dominant else [0.0]
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, kms, ctx.launch_count - launches0

    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    t_wall0 = time.time()
    ms, kms, launches = timed(f9.WINDOWED_SINC, args.steps, args.warmup)
    t_wall1 = time.time()
    clocks = sampler.stop(t_wall0, t_wall1)
    ms_l, kms_l, _ = timed(f9.LAGRANGE, args.steps, args.warmup)
    total_out = out_samples
    if world > 1:
        t = torch.tensor([float(out_samples)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        total_out = float(t.item())

    if args.kernel_only:            # development aid: kernel times only, no JSON contract line
        if rank == 0:
            print(json.dumps({"sinc_kernel_ms": sum(kms) / len(kms), "lagrange_kernel_ms": sum(kms_l) / len(kms_l), "step_ms": ms / args.steps,
                              "frac": alg_dom / (sum(kms) / len(kms) * 1e-3) / 1e9 / peaks["hbm_gbs"] if dominant else None}), flush=True)
        return

    # ---- e2e: file bytes through the batch job flow, H2D + D2H inside the timed region ---------------------------------------
    # 24-bit payloads (what the WAV reader holds, Source/MainComponent.cpp:734-739) in pinned host memory -> f9_process_batch ->
    # 24-bit payloads of the converted files (what the writer takes, :784-801) in pinned host memory, tail-scan results per file.
    e2e = None
    pcie = None
    if payload_h is not None:
        pcie = pcie_probe(torch, dev, barrier, world, dist if world > 1 else None)
        n_outs = [r.n_out for r in rus]
        ob = [r.s.num_ch * n * 3 for r, n in zip(rus, n_outs)]
        ooffs = np.concatenate([[0], np.cumsum([(b + 63) // 64 * 64 for b in ob])]).astype(np.int64)
        out_h = torch.empty(int(ooffs[-1]) + 64, dtype=torch.uint8, pin_memory=True)
        jobs = (f9.Job * len(rus))()
        results = (f9.Result * len(rus))()
        for i, r in enumerate(rus):
            s, j = r.s, jobs[i]
            j.src_pcm, j.src_fmt, j.src_ch = payload_h.data_ptr() + int(offs[i]), f9.PCM_S24LE, s.num_ch
            j.numCh, j.captured_frames = s.num_ch, s.cap_frames
            j.latency_samples, j.original_length = s.latency * s.num_ch, s.src_frames
            j.fs_in, j.fs_out, j.interp_kind = float(s.fs_in), float(s.fs_out), f9.WINDOWED_SINC
            j.flags = f9.JOB_PCM24 | (f9.JOB_TAIL_SCAN if s.tail else 0)
            j.tail_window, j.tail_hop, j.tail_required, j.tail_mode = int(s.fs_in * 0.1), int(s.fs_in * 0.05), 3, f9.TAIL_RMS
            j.has_nf, j.nf_db, j.margin_pct = 1, -90.0, 0.0
            j.out_pcm24 = out_h.data_ptr() + int(ooffs[i])
        h2d = int(sum(nb))
        d2h = int(sum(ob)) + len(rus) * 8

        def e2e_step():
            rc = L.f9_process_batch(ctx.handle, jobs, len(rus), results)
            if rc:
                ctx._check(rc)

        e2e_steps = max(2, min(args.steps, 10))
        for _ in range(3):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        e2e_each = []
        for _ in range(e2e_steps):
            ts = time.perf_counter()
            e2e_step()                                      # blocking: returns when the payloads are in host memory
            e2e_each.append(1e3 * (time.perf_counter() - ts))
        barrier()
        e2e_ms = 1e3 * (time.perf_counter() - t0)
        if world > 1:
            t = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_ms = float(t.item())
        assert all(results[i].status == 0 and results[i].out_frames == n_outs[i] for i in range(len(rus)))
        # the e2e payload of file 0 is the 24-bit packing of exactly what the resident leg produced from the same samples
        step(f9.WINDOWED_SINC); torch.cuda.synchronize(dev)
        r0 = rus[0]
        res0 = torch.stack([d_out[r0.out_at + c * r0.out_stride: r0.out_at + c * r0.out_stride + r0.n_out] for c in range(r0.s.num_ch)])
        q = torch.clamp(torch.round(res0.to(torch.float64) * 2147483647.0), -2147483648.0, 2147483647.0).to(torch.int64) >> 8
        q = q.t().contiguous().reshape(-1)
        want = torch.stack([(q >> (8 * b)) & 0xff for b in range(3)], dim=-1).reshape(-1).to(torch.uint8).cpu()
        got = out_h[int(ooffs[0]): int(ooffs[0]) + ob[0]]
        assert torch.equal(want, got), "resident and e2e legs disagree"
        step_s = e2e_ms / e2e_steps * 1e-3
        e2e = {"value": total_out / step_s / 1e6, "unit": "Msamples/s", "ms_per_step": e2e_ms / e2e_steps,
               "ms_best_step": min(e2e_each), "ms_each_step": [round(v, 2) for v in e2e_each],
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "achieved_h2d_gbs": round(h2d / step_s / 1e9, 2), "achieved_d2h_gbs": round(d2h / step_s / 1e9, 2),
               "frac_of_probe_both": round((h2d + d2h) / step_s / 1e9 / pcie["both_gbs"], 3) if pcie and pcie["both_gbs"] else None,
               "api": "f9_process_batch: f9_job::src_pcm (24-bit interleaved file payload, pinned) -> tail scan, trimLatency, WindowedSinc -> "
                      "F9_JOB_PCM24 payload (pinned); chunks pipelined over upload / kernel / download streams, one host wait per call; per rank"}

    if rank == 0:
        per_step_ms = ms / args.steps
        value = total_out / (per_step_ms * 1e-3) / 1e6
        kavg = sum(kms) / len(kms)
        kavg_l = sum(kms_l) / len(kms_l)
        ach = alg_dom / (kavg * 1e-3) / 1e9 if dominant else 0.0
        ach_l = alg_dom / (kavg_l * 1e-3) / 1e9 if dominant else 0.0
        if max(ach, ach_l) > 1.5 * peaks["hbm_gbs"]:
            raise RuntimeError(f"timed region cannot have contained the work: {ach:.0f} / {ach_l:.0f} GB/s against a "
                               f"{peaks['hbm_gbs']:.0f} GB/s HBM peak")
        sm_mhz = clocks.get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)
        fp32_peak = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12                    # TFLOP/s at the clock seen under load
        dom_out = out_by_group[dominant] if dominant else 0
        fp32_ach = 2.0 * 200 * dom_out / (kavg * 1e-3) / 1e12 if dominant else 0.0
        traffic = TRAFFIC.get((args.workload, len(units))) if (world == 1 or not strong) else None
        s0 = specs[0]
        line = {
            "metric": "resampled output Msamples/s (all channels)", "value": value, "unit": "Msamples/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step_ms,
            "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "interpolator": "WindowedSinc (200 taps)",
                       "files_per_gpu" if not strong else "files_in_workload": n_files, "channels": s0.num_ch,
                       "fs_in": s0.fs_in if not args.workload.startswith("config5") else list(W.MIXED_RATES), "fs_out": s0.fs_out,
                       "seconds_per_file": s0.src_frames / s0.fs_in,
                       "tail_scan": "RMS, 100 ms window / 50 ms hop / 3 consecutive" if tail_rus else "none",
                       "trim": "fused into the resampler", "l2": "inputs (%.2f GB on rank 0) larger than L2" % (in_floats * 4 / 1e9),
                       "parallelism": (f"one workload partitioned over {world} GPUs by f9_multi_partition: {n_units_total} units, {len(units)} on rank 0 (strong, no collective)"
                                       if strong else f"files x{world} (weak, no collective)"),
                       "host_binding": (f"rank bound to the {numa_cpus} CPUs of its GPU's NUMA node" if numa_cpus else "none"),
                       "plan": "created once outside the timed region (tables, tile records, tensor maps); value times f9_dev_tail_scan_batch + f9_resample_plan_run per step; e2e includes planning"},
            "roofline": {"kernel": kernel_name(0, dominant[0], dominant[1]) + f"; {dominant[0]} -> {dominant[1]} Hz; timed with CUDA events on the library's stream around f9_resample_plan_run (the FIR launch and its fp32 redo-check launch; the tile-table launch runs only at the plan's first use)" if dominant else None,
                         "bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"],
                         "peak_source": peak_src, "traffic": traffic[0] if traffic else None, "traffic_source": traffic[1] if traffic else None,
                         "algorithmic_bytes_per_launch": alg_dom, "kernel_ms": kavg,
                         "note": "200 taps = 400 FLOP per output: above the FP32 ridge on CUDA cores, so the taps run on the tensor "
                                 "cores and the stage is measured against the HBM roofline it is meant to reach",
                         "fp32_equivalent": {"achieved_tflops": fp32_ach, "cuda_core_peak_tflops": fp32_peak, "frac": fp32_ach / fp32_peak,
                                             "note": "useful FLOP (400 per output) against the FP32 CUDA-core peak at the SM clock seen: "
                                                     "what a CUDA-core FIR could reach at most"},
                         "tensor": {"useful_tflops": fp32_ach, "peak_tflops": peaks.get("bf16_tflops"),
                                    "frac": fp32_ach / peaks["bf16_tflops"] if peaks.get("bf16_tflops") else None,
                                    "note": "issued MMA work is ~4.5x the useful FLOP (3 fp16 products per tap, band padding)"}},
            "lagrange": {"value": total_out / (ms_l / args.steps * 1e-3) / 1e6, "unit": "Msamples/s", "ms_per_step": ms_l / args.steps,
                         "roofline": {"kernel": kernel_name(1, dominant[0], dominant[1]) if dominant else None, "bound": "hbm", "achieved": ach_l,
                                      "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach_l / peaks["hbm_gbs"], "kernel_ms": kavg_l}},
            "e2e": e2e if e2e else {"value": None, "unit": "Msamples/s", "h2d_bytes_per_step": None, "d2h_bytes_per_step": None,
                                    "note": "the file-bytes leg runs on workloads whose units are whole files (configs 1, 2, 5); --no-e2e skips it"},
            "pcie_probe": pcie,
            "gpu_launches": launches,
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu:
            cores = os.cpu_count() or 1
            nref = max(1, min(args.ref_files, W.CONFIGS[args.workload][4]))
            n, dt = cpu_leg(args.workload, 0, nref, cores)
            line["cpu_baseline"] = {"value": n / dt / 1e6, "unit": "Msamples/s", "cores": cores, "kind": "port",
                                    "sample": f"{nref} of {W.CONFIGS[args.workload][4]} files (tail scan + trimLatency + WindowedSinc), oracle port, {dt:.1f} s"}
            n1f = max(1, min(nref, 8))
            n1, dt1 = cpu_leg(args.workload, 0, n1f, 1)
            line["cpu_baseline_1_thread"] = {"value": n1 / dt1 / 1e6, "unit": "Msamples/s", "cores": 1, "kind": "port",
                                             "sample": f"{n1f} files on one thread, {dt1:.1f} s"}
        emit(line)
    for p in plans.values():
        L.f9_plan_destroy(p)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=W.DEFAULT, choices=sorted(W.CONFIGS))
    ap.add_argument("--scaling", default="auto", choices=["auto", "weak", "strong"], help="config 2: weak by default (files per GPU fixed); strong shards one workload")
    ap.add_argument("--files", type=int, default=None, help="files per GPU (weak) or in the workload (strong); default: the config's")
    ap.add_argument("--seconds", type=float, default=None, help="source seconds per file (default: the config's)")
    ap.add_argument("--ref-files", type=int, default=256, help="files in the CPU baseline sample (256 = the whole config-2 workload, ~5 s on 16 cores)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
    ap.add_argument("--no-e2e", action="store_true", help="skip the file-bytes e2e leg and the PCIe probe")
    ap.add_argument("--kernel-only", action="store_true", help="development: print kernel times only")
    ap.add_argument("--unaligned", action="store_true", help="resident captures start (not their trimmed start) on 16 bytes")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl != "reference":
        args.warmup = 3
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if not args.kernel_only:
        claim_stdout()
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_gpu(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
