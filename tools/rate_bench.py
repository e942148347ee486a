import sys, os, ctypes as C, json
sys.path.insert(0, "/root/repo")
import torch
import __graft_entry__ as g
f9 = g._load_pkg()
if os.environ.get("F9DSP_DIAG_LIB"):          # development: the -DF9_DIAG build (make -C f9-juce-resampler-studio_b200 DIAG=1)
    f9.LIB_PATH = os.path.join(os.path.dirname(os.path.dirname(f9.LIB_PATH)), "lib_diag", "libf9dsp.so")
dev = torch.device("cuda", 0)
ctx = f9.Context(0); stream = torch.cuda.Stream(dev); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
L = f9.lib()
for kv in os.environ.get("OPTS", "").split(","):          # e.g. OPTS=F9_UMMA_STAGES=4,F9_UMMA_NOCTA2=1
    if kv:
        ctx.set_option(kv.split("=")[0], int(kv.split("=")[1]) if "=" in kv else 1)
gen = torch.Generator(device=dev); gen.manual_seed(1)
def plan_time(kind, fs_in, fs_out, nch, n_in):
    x = torch.randn((nch, n_in), generator=gen, device=dev, dtype=torch.float32) * 0.25
    no = f9.resampled_length(n_in, fs_in, fs_out)
    y = torch.empty((nch, no), dtype=torch.float32, device=dev)
    segs = (f9.ResampleSeg * nch)(*[f9.ResampleSeg(x[c].data_ptr(), 0, n_in, y[c].data_ptr(), 0, no) for c in range(nch)])
    plan = C.c_void_p(None)
    ctx._check(L.f9_resample_plan_create(ctx.handle, kind, fs_in / fs_out, segs, nch, C.byref(plan)))
    for _ in range(3): ctx._check(L.f9_resample_plan_run(plan))
    ts = []
    for _ in range(7):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); ctx._check(L.f9_resample_plan_run(plan)); e1.record(stream); e1.synchronize(); ts.append(e0.elapsed_time(e1))
    L.f9_plan_destroy(plan)
    ts.sort(); ms = ts[3]
    print(kind, fs_in, fs_out, round(ms, 4), "ms", round(4.0 * nch * (n_in + no) / ms / 1e6 / 6551.4 * 100, 1), "%", flush=True)
# usage: rate_bench.py [kinds, e.g. 0,1] [fs_in:fs_out, ...]   (512 channels of 10 s each)
kinds = [int(k) for k in sys.argv[1].split(",")] if len(sys.argv) > 1 else [0, 1]
rates = [tuple(int(v) for v in a.split(":")) for a in sys.argv[2:]] or [(44100, 48000), (88200, 48000)]
for kind in kinds:
    for fs_in, fs_out in rates:
        plan_time(kind, fs_in, fs_out, 512, 10 * fs_in)
