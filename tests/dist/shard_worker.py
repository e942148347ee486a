"""Worker of tests/test_sharding_cpu.py: run under torchrun with the gloo backend (CPU), one process per "GPU".
Each rank takes its shard of a job list (f9_shard_units of the product: SURVEY 8(e), no data-path collective), computes the shard's
scalar results with the oracle standing in for the device, and the results are gathered on rank 0 -- the same host-side
flow bench.py and a multi-GPU host use around the library.  Rank 0 writes the gathered results as JSON to argv[1]."""
import importlib.util
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

spec = importlib.util.spec_from_file_location("f9dsp_w", os.path.join(ROOT, "f9-juce-resampler-studio_b200", "py", "f9dsp.py"))
F9 = importlib.util.module_from_spec(spec)
sys.modules["f9dsp_w"] = F9
spec.loader.exec_module(F9)


def shard_units(costs, world):
    """The product's greedy packing (f9_shard_units, host code of libf9dsp.so): bins of unit indices."""
    import ctypes as C
    n = len(costs)
    bins = (C.c_int * n)()
    assert F9.lib().f9_shard_units((C.c_longlong * n)(*costs), n, world, bins) == 0
    return [[i for i in range(n) if bins[i] == r] for r in range(world)]


def job(i):
    rng = np.random.default_rng(100 + i)
    frames = 3000 + 500 * (i % 5)
    cap = (rng.standard_normal((2, frames)) * 1e-4).astype(np.float32)
    lat = 17 * i + 3
    cap[:, lat] += 0.9
    return cap, lat, frames


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    n_jobs = 11
    costs = [job(i)[2] for i in range(n_jobs)]
    mine = shard_units(costs, world)[rank]
    out = {}
    for i in mine:
        cap, lat, frames = job(i)
        peak = O.find_peak_position(cap, 0.1)
        trimmed, copied = O.trim_latency(cap, 2 * peak, 1000)
        out[i] = [int(peak), int(copied), float(O.calculate_rms(trimmed))]
    gathered = [None] * world
    dist.gather_object(out, gathered if rank == 0 else None, dst=0)
    # timing reduction as in bench.py: the step time of the job is the max over ranks
    t = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        merged = {}
        for g in gathered:
            merged.update({str(k): v for k, v in g.items()})
        json.dump({"results": merged, "max_ms": float(t.item()), "world": world, "shards": shard_units(costs, world)}, open(sys.argv[1], "w"))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
