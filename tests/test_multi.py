"""The batch job flow over several GPUs (include/f9dsp.h section H).

CPU: the partition itself (host code in the product): greedy packing, channel groups and time segments with halos cover every
(channel, output) of a split job exactly once.
GPU: f9_multi with two contexts and two host threads (on one GPU when the box has one: the device list may repeat an ordinal) gives
the single-context results; concurrent first use of every kernel family from two threads (the process-wide state is call_once /
atomic); on a box with >= 2 GPUs the same over distinct devices."""
import ctypes as C
import threading

import numpy as np
import pytest

TOL = 2.0 ** -20


def _dummy_job(f9, J, i, num_ch, frames, src, fs_in, fs_out, flags=0, lat=0):
    """Descriptor only (pointers are not dereferenced by the partition)."""
    keep = (C.POINTER(C.c_float) * num_ch)()
    for c in range(num_ch):
        keep[c] = C.cast(0x1000, C.POINTER(C.c_float))
    J[i].captured, J[i].numCh, J[i].captured_frames = keep, num_ch, frames
    J[i].latency_samples, J[i].original_length = lat * num_ch, src
    J[i].fs_in, J[i].fs_out, J[i].interp_kind, J[i].flags = fs_in, fs_out, 0, flags
    J[i].tail_window, J[i].tail_hop, J[i].tail_required = 4800, 2400, 3
    J[i].out = keep
    J[i].out_capacity = 1 << 30
    return keep


def test_shard_units_is_a_balanced_partition(f9):
    rng = np.random.default_rng(0)
    for world in (1, 2, 4, 8):
        costs = (C.c_longlong * 97)(*[int(c) for c in rng.integers(1000, 500000, 97)])
        bins = (C.c_int * 97)()
        assert f9.lib().f9_shard_units(costs, 97, world, bins) == 0
        loads = [sum(costs[i] for i in range(97) if bins[i] == r) for r in range(world)]
        assert sorted(set(bins)) == list(range(world))
        assert max(loads) - min(loads) <= max(costs)                                   # greedy LPT bound


def test_partition_config3_shape(f9):
    """Config 3: one 64-channel 10-minute file 48 k -> 192 k over 8 GPUs: channel groups x time segments, each (channel, output)
    in exactly one unit, loads within a few percent, the tail scan as a unit of its own."""
    J = (f9.Job * 1)()
    keep = _dummy_job(f9, J, 0, 64, 28_800_000 + 5000, 28_800_000, 48000.0, 192000.0, flags=f9.JOB_TAIL_SCAN, lat=100)
    n_out = f9.resampled_length(28_800_000, 48000.0, 192000.0)
    for world in (2, 4, 8):
        U, nu = f9.partition(J, 1, world)
        cover = {}
        tails = 0
        for k in range(nu):
            u = U[k]
            assert u.job == 0 and 0 <= u.device < world
            if u.tail_only:
                tails += 1
                continue
            assert u.num_out > 0 and u.num_ch > 0
            for c in range(u.ch0, u.ch0 + u.num_ch):
                cover.setdefault(c, []).append((u.n0, u.num_out))
        assert tails == 1 and sorted(cover) == list(range(64))
        for c, segs in cover.items():
            segs.sort()
            assert segs[0][0] == 0 and all(a[0] + a[1] == b[0] for a, b in zip(segs, segs[1:])) and segs[-1][0] + segs[-1][1] == n_out
        loads = [sum(U[k].cost for k in range(nu) if U[k].device == r and not U[k].tail_only) for r in range(world)]
        assert max(loads) <= 1.05 * min(loads), loads
    del keep


def test_partition_config5_shape(f9):
    """Config 5: 4096 mixed-rate stereo files: files are the units (none is large against a GPU's share), packed by output count."""
    rates = [44100.0, 48000.0, 88200.0, 96000.0, 192000.0]
    n = 4096
    J = (f9.Job * n)()
    keeps = [_dummy_job(f9, J, i, 2, int(10 * rates[i % 5]) + 4096, int(10 * rates[i % 5]), rates[i % 5], 48000.0) for i in range(n)]
    U, nu = f9.partition(J, n, 8)
    assert nu == n and all(U[k].num_out == 0 and U[k].num_ch == 2 and not U[k].tail_only for k in range(nu))
    loads = [sum(U[k].cost for k in range(nu) if U[k].device == r) for r in range(8)]
    assert max(loads) - min(loads) <= 2 * 480000
    del keeps


def _jobs(O, f9, rng):
    """Small files of mixed rates and kinds + one long multichannel file that the partition must split."""
    jobs = []
    for i in range(7):
        fs_in = [44100, 96000, 48000, 88200][i % 4]
        fs_out = [48000, 44100, 192000, 48000][i % 4]
        src, lat = 9000 + 500 * i, 33 * i + 5
        cap = (rng.uniform(-0.5, 0.5, (2, O.recording_length(src, lat) + 6000))).astype(np.float32)
        cap[:, src + lat + 2000:] *= 1e-6
        jobs.append(dict(captured=cap, latency_samples=2 * lat, original_length=src, fs_in=fs_in, fs_out=fs_out, kind=i % 2,
                         tail=(2048, 1024, 3, 0, True, -90.0, 0.0), pcm24=(i % 3 == 0)))
    src, lat = 700000, 77
    big = (rng.uniform(-0.5, 0.5, (6, O.recording_length(src, lat) + 20000))).astype(np.float32)
    big[:, src + lat + 5000:] *= 1e-6
    jobs.append(dict(captured=big, latency_samples=6 * lat, original_length=src, fs_in=48000, fs_out=96000, kind=0,
                     tail=(4800, 2400, 3, 0, True, -90.0, 0.0)))
    bigl = (rng.uniform(-0.5, 0.5, (2, 900000))).astype(np.float32)
    jobs.append(dict(captured=bigl, latency_samples=0, original_length=900000, fs_in=96000, fs_out=44100, kind=1, pcm24=True))
    jobs.append(dict(captured=bigl[:, :500000], latency_samples=10, original_length=400000, fs_in=96000, fs_out=44100, kind=0, remove_dc=True))
    return jobs


def _compare(single, multi):
    (o1, p1, r1), (o2, p2, r2) = single, multi
    for i in range(len(r1)):
        a = {k: v for k, v in r1[i].items()}
        b = {k: v for k, v in r2[i].items() if k != "device"}
        assert a == b, (i, a, b)
        assert o1[i].shape == o2[i].shape
        # same kernels, same arithmetic per output whatever the tiling: pieces reproduce the whole within rounding of the tile grid
        assert np.max(np.abs(o1[i] - o2[i])) <= TOL / 4, i
        if p1[i] is not None:
            d = np.abs(p1[i].astype(np.int32) - p2[i].astype(np.int32))
            assert p1[i].shape == p2[i].shape and np.mean(d != 0) < 1e-3, i


@pytest.mark.gpu
def test_multi_two_contexts_match_single(ctx, O, f9):
    jobs = _jobs(O, f9, np.random.default_rng(21))
    single = ctx.process_batch(jobs)
    assert all(r["status"] == 0 for r in single[2])
    m = f9.Multi([0, 0])
    try:
        multi = m.process_batch(jobs)
        _compare(single, multi)
        assert {r["device"] for r in multi[2]} == {0, 1}
        # and against the oracle for the split job
        big = jobs[7]
        trimmed, _ = O.trim_latency(big["captured"], big["latency_samples"], big["original_length"])
        ref = O.resample_channel(0, 0.5, trimmed[3], multi[0][7].shape[1])[0]
        assert np.max(np.abs(multi[0][7][3] - ref)) <= TOL
        stop, _ = O.tail_scan(big["captured"], big["original_length"] + 77, 4800, 2400, 3, 0, True, -90.0, 0.0)
        assert multi[2][7]["tail_stop_frame"] == stop
    finally:
        m.close()


@pytest.mark.gpu
def test_multi_distinct_devices(O, f9):
    if f9.device_count() < 2:
        pytest.skip("one GPU on this box (the two-context test covers the host logic)")
    jobs = _jobs(O, f9, np.random.default_rng(22))
    c = f9.Context(0)
    try:
        single = c.process_batch(jobs)
    finally:
        c.close()
    m = f9.Multi(list(range(min(f9.device_count(), 8))))
    try:
        _compare(single, m.process_batch(jobs))
    finally:
        m.close()


@pytest.mark.gpu
def test_concurrent_first_use_from_two_threads(O, f9):
    """Two host threads, each with a fresh context, hit every kernel family for the first time at once (driver entry points,
    function attributes, table builds): results are the single-threaded ones."""
    rng = np.random.default_rng(23)
    x = rng.uniform(-0.5, 0.5, (2, 60000)).astype(np.float32)
    cases = [(96000, 44100, 0), (44100, 48000, 0), (48000, 192000, 0), (96000, 44100, 1), (88200, 48000, 0), (48000, 96000, 1)]
    ref = {}
    c0 = f9.Context(0)
    try:
        for cs in cases:
            ref[cs] = c0.resample(x, *cs)
    finally:
        c0.close()
    out = [dict(), dict()]
    errs = []

    def work(t):
        try:
            c = f9.Context(0)
            try:
                for cs in (cases if t == 0 else cases[::-1]):
                    out[t][cs] = c.resample(x, *cs)
                    assert c.find_peak_position(x, 0.1) == O.find_peak_position(x, 0.1)
            finally:
                c.close()
        except Exception as e:      # noqa: BLE001
            errs.append(e)

    for rep in range(3):
        ts = [threading.Thread(target=work, args=(t,)) for t in range(2)]
        [t.start() for t in ts]
        [t.join() for t in ts]
        assert not errs, errs
        for t in range(2):
            for cs in cases:
                assert np.array_equal(out[t][cs], ref[cs]), (rep, t, cs)
