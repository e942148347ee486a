#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on BASELINE.json's config.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME] [--files F]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...          (N > 1, one rank per GPU)

metric  : resampled output Msamples/s, all channels (BASELINE.json "metric")
workload: configs[1] -- batch of 256 stereo files 96 kHz -> 44.1 kHz with latency trim and tail-silence detection,
          10 s per file, converted with the WindowedSinc interpolator (the Lagrange numbers ride along in "lagrange").
step    : one pass of the hot path over the whole batch: reverb-tail scan of every capture + trimLatency fused into the
          polyphase resampler, captures resident in HBM (value) or in pinned host memory through f9_process_batch (e2e).
N > 1   : weak scaling; every rank owns its own batch (files rank*F .. rank*F+F-1), no data-path collective;
          time = max over ranks, value = all ranks' output samples / that time.
--impl reference : the reference's CPU implementation of the path.  JUCE is not vendored by the reference and the
          reference cannot be compiled here, so this times the oracle port (oracle/, scalar C++, one interpolator
          object per channel as JUCE runs it) on all host cores, on a bounded sample of the same workload.
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
def cpu_leg(batch, kind: int, files: int, threads: int):
    """One bounded sample of the workload on the host cores: tail scan + trimLatency + conversion per file, files
    statically partitioned over `threads` (ctypes calls release the GIL).  Returns (out_samples, seconds)."""
    from oracle import oracle as O

    caps = W.fill_host_numpy(batch, 0, files)
    ratio = batch.fs_in / batch.fs_out
    n_out = -((-batch.src_frames * batch.fs_out) // batch.fs_in)
    win, hop = int(batch.fs_in * 0.1), int(batch.fs_in * 0.05)
    O.lib()

    def work(f0, f1):
        for i in range(f0, f1):
            lat = W.latency_of(i)
            O.tail_scan(caps[i], batch.src_frames + lat, win, hop, 3, 0, True, -90.0, 0.0)
            trimmed, _ = O.trim_latency(caps[i], lat * batch.num_ch, batch.src_frames)
            O.resample_channels(kind, ratio, trimmed, n_out, threads=1)

    threads = max(1, min(threads, files))
    bounds = [files * k // threads for k in range(threads + 1)]
    ts = [threading.Thread(target=work, args=(bounds[k], bounds[k + 1])) for k in range(threads)]
    t0 = time.perf_counter()
    [t.start() for t in ts]
    [t.join() for t in ts]
    dt = time.perf_counter() - t0
    return files * batch.num_ch * n_out, dt


def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    batch = W.describe(args.workload, args.files)
    cores = os.cpu_count() or 1
    files = max(cores, min(batch.files, args.ref_files))
    for _ in range(args.warmup):
        cpu_leg(batch, 0, min(files, cores), cores)
    total, secs = 0, 0.0
    for _ in range(args.steps):
        n, dt = cpu_leg(batch, 0, files, cores)
        total += n; secs += dt
    v = total / secs / 1e6
    line = {"impl": "reference", "metric": "resampled output Msamples/s (all channels)", "value": v, "unit": "Msamples/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "interpolator": "WindowedSinc", "files_in_workload": batch.files,
                       "fs_in": batch.fs_in, "fs_out": batch.fs_out, "channels": batch.num_ch, "seconds_per_file": batch.src_frames / batch.fs_in},
            "cpu_baseline": {"value": v, "unit": "Msamples/s", "cores": cores, "kind": "port",
                             "sample": f"{files} of {batch.files} files per step (tail scan + trimLatency + WindowedSinc), oracle port, g++ -O2 -ffp-contract=off"},
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


# dram__bytes_read.sum + dram__bytes_write.sum of one umma_fir_kernel launch (ncu --set full, profiles/), by files per GPU
TRAFFIC_BYTES_PER_LAUNCH = {256: 2.054042e9 + 873.202944e6}    # profiles/r01_v6_umma_fir_full.txt (dram read + write of one launch)


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

    nfiles = args.files or W.CONFIGS[args.workload][4]
    first_file = rank * nfiles
    batch = W.describe(args.workload, nfiles, first_file=first_file)
    ctx = f9.Context(local_rank)                      # raises when the CUDA library or the GPU is missing
    # One explicit stream carries the library's kernels AND the timing events (the legacy default stream has
    # handle 0, which f9_set_stream reads as "use the context's own stream": events there would time nothing).
    stream = torch.cuda.Stream(dev)
    assert stream.cuda_stream != 0
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    L = f9.lib()

    caps0 = W.fill_device(batch, first_file, dev)                         # [files, ch, cap] resident in HBM
    # Resident layout = the library's own upload layout (f9_batch.cu): each capture sits so that its TRIMMED start
    # (capture + latency frames) is 16-byte aligned; --unaligned keeps the capture's first frame aligned instead.
    pads = [0 if args.unaligned else (-lat) % 4 for lat in batch.latency_frames]
    cap_stride = batch.cap_frames + 64
    caps = torch.zeros((batch.files, batch.num_ch, cap_stride), dtype=torch.float32, device=dev)
    for i in range(batch.files):
        caps[i, :, pads[i]:pads[i] + batch.cap_frames] = caps0[i]
    del caps0
    n_out = f9.resampled_length(batch.src_frames, batch.fs_in, batch.fs_out)
    out_stride = (n_out + 63) // 64 * 64
    outs = torch.empty((batch.files, batch.num_ch, out_stride), dtype=torch.float32, device=dev)
    ratio = batch.fs_in / batch.fs_out
    win, hop, req = int(batch.fs_in * 0.1), int(batch.fs_in * 0.05), 3
    max_polls = max(1, (batch.cap_frames - batch.src_frames) // hop)
    stops = torch.empty(batch.files, dtype=torch.int64, device=dev)
    flags = torch.empty(batch.files * max_polls, dtype=torch.int32, device=dev)

    # descriptors: tail scan over the captures, resample segments with trimLatency fused as a pointer offset
    bufs = (f9.DevBuffer * batch.files)()
    tails = (f9.TailParams * batch.files)()
    segs = (f9.ResampleSeg * (batch.files * batch.num_ch))()
    for i in range(batch.files):
        base = caps.data_ptr() + 4 * (i * batch.num_ch * cap_stride + pads[i])
        bufs[i] = f9.DevBuffer(base, cap_stride, batch.num_ch, batch.cap_frames)
        lat = batch.latency_frames[i]
        tails[i] = f9.TailParams(batch.src_frames + lat, win, hop, req, f9.TAIL_RMS, 1, -90.0, 0.0)
        copied = max(0, min(batch.src_frames, batch.cap_frames - lat))    # trimLatency arithmetic (MainComponent.cpp:833-845)
        for c in range(batch.num_ch):
            segs[i * batch.num_ch + c] = f9.ResampleSeg(base + 4 * (c * cap_stride + lat), 0, copied,
                                                        outs.data_ptr() + 4 * (i * batch.num_ch + c) * out_stride, 0, n_out)
    plans = {}
    for kind in (f9.WINDOWED_SINC, f9.LAGRANGE):
        p = C.c_void_p(None)
        ctx._check(L.f9_resample_plan_create(ctx.handle, kind, ratio, segs, len(segs), C.byref(p)))
        plans[kind] = p

    def step(kind):
        ctx._check(L.f9_dev_tail_scan_batch(ctx.handle, bufs, tails, batch.files, stops.data_ptr(), flags.data_ptr(), max_polls))
        ctx._check(L.f9_resample_plan_run(plans[kind]))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    out_samples = batch.files * batch.num_ch * n_out
    alg_bytes = (4.0 + 4.0 * ratio) * out_samples                         # 4 + 4*ratio per output sample

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
            ctx._check(L.f9_dev_tail_scan_batch(ctx.handle, bufs, tails, batch.files, stops.data_ptr(), flags.data_ptr(), max_polls))
            k0[s].record(stream)
            ctx._check(L.f9_resample_plan_run(plans[kind]))
            k1[s].record(stream)
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        kms = [a.elapsed_time(b) for a, b in zip(k0, k1)]
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

    if args.kernel_only:            # development aid: kernel times only, no JSON contract line
        if rank == 0:
            print(json.dumps({"cfg": os.environ.get("F9_BANDED_CFG", ""), "sinc_kernel_ms": sum(kms) / len(kms),
                              "lagrange_kernel_ms": sum(kms_l) / len(kms_l), "step_ms": ms / args.steps}), flush=True)
        return

    # ---- e2e: host buffers through f9_process_batch, H2D + D2H inside the timed region ----
    caps_h = torch.empty((batch.files, batch.num_ch, batch.cap_frames), dtype=torch.float32, pin_memory=True)
    for i in range(batch.files):                                          # host buffers are plain juce::AudioBuffer-style planes
        caps_h[i].copy_(caps[i, :, pads[i]:pads[i] + batch.cap_frames])
    outs_h = torch.empty((batch.files, batch.num_ch, n_out), dtype=torch.float32, pin_memory=True)
    torch.cuda.synchronize(dev)
    fp = C.POINTER(C.c_float)
    chan_in = (fp * (batch.files * batch.num_ch))()
    chan_out = (fp * (batch.files * batch.num_ch))()
    jobs = (f9.Job * batch.files)()
    results = (f9.Result * batch.files)()
    for i in range(batch.files):
        for c in range(batch.num_ch):
            chan_in[i * batch.num_ch + c] = C.cast(caps_h.data_ptr() + 4 * (i * batch.num_ch + c) * batch.cap_frames, fp)
            chan_out[i * batch.num_ch + c] = C.cast(outs_h.data_ptr() + 4 * (i * batch.num_ch + c) * n_out, fp)
        j = jobs[i]
        j.captured = C.cast(C.byref(chan_in, C.sizeof(fp) * i * batch.num_ch), C.POINTER(fp))
        j.out = C.cast(C.byref(chan_out, C.sizeof(fp) * i * batch.num_ch), C.POINTER(fp))
        j.numCh, j.captured_frames = batch.num_ch, batch.cap_frames
        j.latency_samples, j.original_length = batch.latency_frames[i] * batch.num_ch, batch.src_frames
        j.fs_in, j.fs_out, j.interp_kind = float(batch.fs_in), float(batch.fs_out), f9.WINDOWED_SINC
        j.flags = f9.JOB_TAIL_SCAN
        j.tail_window, j.tail_hop, j.tail_required, j.tail_mode = win, hop, req, f9.TAIL_RMS
        j.has_nf, j.nf_db, j.margin_pct = 1, -90.0, 0.0
        j.out_capacity = n_out
    h2d = batch.files * batch.num_ch * batch.cap_frames * 4
    d2h = batch.files * batch.num_ch * n_out * 4 + batch.files * 8

    def e2e_step():
        rc = L.f9_process_batch(ctx.handle, jobs, batch.files, results)
        if rc:
            ctx._check(rc)

    e2e_steps = max(2, min(args.steps, 5))
    for _ in range(3):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    e2e_each = []
    for _ in range(e2e_steps):
        ts = time.perf_counter()
        e2e_step()                                      # blocking: returns when the outputs are in host memory
        e2e_each.append(1e3 * (time.perf_counter() - ts))
    e2e_best = min(e2e_each)
    barrier()
    e2e_ms = 1e3 * (time.perf_counter() - t0)
    if world > 1:
        t = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    assert all(results[i].status == 0 and results[i].out_frames == n_out for i in range(batch.files))
    # the e2e outputs are the same samples the resident leg produced
    step(f9.WINDOWED_SINC); torch.cuda.synchronize(dev)
    chk = float((outs[0, :, :n_out].cpu() - outs_h[0]).abs().max())
    assert chk == 0.0, f"resident and e2e legs disagree: {chk}"

    if rank == 0:
        per_step_ms = ms / args.steps
        value = world * out_samples / (per_step_ms * 1e-3) / 1e6
        kavg = sum(kms) / len(kms)
        kavg_l = sum(kms_l) / len(kms_l)
        ach = alg_bytes / (kavg * 1e-3) / 1e9
        ach_l = alg_bytes / (kavg_l * 1e-3) / 1e9
        if max(ach, ach_l) > 1.5 * peaks["hbm_gbs"]:
            raise RuntimeError(f"timed region cannot have contained the work: {ach:.0f} / {ach_l:.0f} GB/s against a "
                               f"{peaks['hbm_gbs']:.0f} GB/s HBM peak")
        sm_mhz = clocks.get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)
        fp32_peak = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12                    # TFLOP/s at the clock seen under load
        fp32_ach = 2.0 * 200 * out_samples / (kavg * 1e-3) / 1e12
        line = {
            "metric": "resampled output Msamples/s (all channels)", "value": value, "unit": "Msamples/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step_ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "interpolator": "WindowedSinc (200 taps, polyphase 147 phases)",
                       "files_per_gpu": batch.files, "channels": batch.num_ch, "fs_in": batch.fs_in, "fs_out": batch.fs_out,
                       "seconds_per_file": batch.src_frames / batch.fs_in, "tail_scan": "RMS, 100 ms window / 50 ms hop / 3 consecutive",
                       "trim": "fused into the resampler", "l2": "inputs (%.2f GB per GPU) larger than L2" % (h2d / 1e9),
                       "parallelism": f"files x{world} (weak, no collective)",
                       "host_binding": (f"rank bound to the {numa_cpus} CPUs of its GPU's NUMA node" if numa_cpus else "none")},
            "roofline": {"kernel": "umma_fir_kernel (WindowedSinc polyphase FIR on tcgen05: TMA-fed CTA pairs, fp16 2-split, fp32 TMEM accumulators; "
                                   "timed with its tile-table and redo-check launches)",
                         "bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"],
                         "peak_source": peak_src, "traffic": TRAFFIC_BYTES_PER_LAUNCH.get(batch.files),
                         "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": kavg,
                         "note": "200 taps = 400 FLOP per output: above the FP32 ridge on CUDA cores, so the taps run on the tensor "
                                 "cores and the stage is measured against the HBM roofline it is meant to reach",
                         "fp32_equivalent": {"achieved_tflops": fp32_ach, "cuda_core_peak_tflops": fp32_peak, "frac": fp32_ach / fp32_peak,
                                             "note": "useful FLOP (400 per output) against the FP32 CUDA-core peak at the SM clock seen: "
                                                     "what a CUDA-core FIR could reach at most"},
                         "tensor": {"useful_tflops": fp32_ach, "peak_tflops": peaks.get("bf16_tflops"),
                                    "frac": fp32_ach / peaks["bf16_tflops"] if peaks.get("bf16_tflops") else None,
                                    "note": "issued MMA work is ~4.5x the useful FLOP (3 fp16 products per tap, band padding)"}},
            "lagrange": {"value": world * out_samples / (ms_l / args.steps * 1e-3) / 1e6, "unit": "Msamples/s", "ms_per_step": ms_l / args.steps,
                         "roofline": {"kernel": "short_kernel (Lagrange polyphase FIR on CUDA cores, fp32: cp.async-staged tiles, slot weights in registers)", "bound": "hbm", "achieved": ach_l,
                                      "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach_l / peaks["hbm_gbs"], "kernel_ms": kavg_l}},
            "e2e": {"value": world * out_samples / (e2e_ms / e2e_steps * 1e-3) / 1e6, "unit": "Msamples/s", "ms_per_step": e2e_ms / e2e_steps,
                    "ms_best_step": e2e_best, "ms_each_step": [round(v, 2) for v in e2e_each],
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "f9_process_batch (pinned host buffers; chunks pipelined over two streams)"},
            "gpu_launches": launches,
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu:
            cores = os.cpu_count() or 1
            n, dt = cpu_leg(batch, 0, args.ref_files, cores)
            line["cpu_baseline"] = {"value": n / dt / 1e6, "unit": "Msamples/s", "cores": cores, "kind": "port",
                                    "sample": f"{args.ref_files} of {batch.files} files (tail scan + trimLatency + WindowedSinc), oracle port, {dt:.1f} s"}
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
    ap.add_argument("--files", type=int, default=None, help="files per GPU (default: the config's)")
    ap.add_argument("--ref-files", type=int, default=256, help="files in the CPU baseline sample (256 = the whole workload, ~10 s on 16 cores)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--kernel-only", action="store_true", help="development: print kernel times only")
    ap.add_argument("--unaligned", action="store_true", help="resident captures start (not their trimmed start) on 16 bytes")
    args = ap.parse_args()
    if args.warmup < 3:
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
