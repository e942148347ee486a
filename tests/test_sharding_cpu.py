"""CPU: the N > 1 path.  The DSP path shards by file / channel / time segment with no data-path collective;
what runs across ranks is host logic: unit packing (f9_shard_units / f9_multi_partition: tests/test_multi.py), halo windows of time
segments, gathering results on rank 0, and the max-over-ranks timing of bench.py.  Covered here with world_size 2 over the gloo backend (no GPU)."""
import importlib.util
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(nproc, script_args, timeout=300):
    env = dict(os.environ, OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", "29533"] + script_args
    return subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)


def test_time_segments_and_halos(f9):
    """Long channels split into output ranges; each range's input window carries its own halo (199 inputs for WindowedSinc,
    4 for Lagrange) and the windows of neighbouring ranges overlap by exactly that."""
    for kind, memory in ((0, 200), (1, 5)):
        prev_last = None
        for n0 in range(0, 1_000_000, 130_001):
            cnt = min(130_001, 1_000_000 - n0)
            first, last = f9.segment_input_range(kind, 48000 / 192000, n0, cnt)
            newest_first = (n0 * 1) // 4 + 1                                             # inputs consumed before / by the first output
            assert last - first >= cnt // 4 and first <= newest_first
            if prev_last is not None:
                assert prev_last - first >= memory - 1                                  # the halo reaches back one memory length
            prev_last = last


def test_world_size_2_gloo_shards_and_gathers(tmp_path, O):
    out = tmp_path / "gathered.json"
    r = _torchrun(2, [os.path.join(ROOT, "tests", "dist", "shard_worker.py"), str(out)])
    assert r.returncode == 0, r.stderr[-2000:]
    got = json.load(open(out))
    assert got["world"] == 2 and got["max_ms"] == 11.0                                  # max over ranks of (10 + rank)
    assert sorted(i for b in got["shards"] for i in b) == list(range(11))
    # the union of the shards' results equals the serial run
    sys.path.insert(0, os.path.join(ROOT, "tests", "dist"))
    import shard_worker as SW
    for i in range(11):
        cap, lat, frames = SW.job(i)
        peak = O.find_peak_position(cap, 0.1)
        trimmed, copied = O.trim_latency(cap, 2 * peak, 1000)
        assert got["results"][str(i)] == [int(peak), int(copied), float(O.calculate_rms(trimmed))]
        assert peak == lat


def test_reference_arm_under_torchrun_prints_one_line_from_rank_0():
    """bench.py --impl reference at N = 2: rank 0 alone times the CPU port and prints the JSON line, rank 1 exits 0."""
    r = _torchrun(2, [os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--ref-files", "1",
                      "--workload", "config1_60s_stereo_44k1_to_48k"], timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["value"] > 0 and d["cpu_baseline"]["kind"] == "port"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
