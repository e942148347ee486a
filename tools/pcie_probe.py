import torch, time
n = 512*1024*1024
h = torch.empty(n, dtype=torch.float32, pin_memory=True); d = torch.empty(n, dtype=torch.float32, device="cuda")
h2 = torch.empty(n//2, dtype=torch.float32, pin_memory=True); d2 = torch.empty(n//2, dtype=torch.float32, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for _ in range(2): d.copy_(h, non_blocking=True); torch.cuda.synchronize()
t=time.perf_counter(); d.copy_(h, non_blocking=True); torch.cuda.synchronize(); dt=time.perf_counter()-t; print("H2D GB/s", n*4/dt/1e9)
t=time.perf_counter(); h2.copy_(d2, non_blocking=True); torch.cuda.synchronize(); dt=time.perf_counter()-t; print("D2H GB/s", n*2/dt/1e9)
t=time.perf_counter()
with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize(); dt=time.perf_counter()-t; print("both: H2D 2GB + D2H 1GB in ms", dt*1e3, "sum GB/s", n*6/dt/1e9)
# many small copies
chunks = h.view(512, -1)
t=time.perf_counter()
for i in range(512): d.view(512,-1)[i].copy_(chunks[i], non_blocking=True)
torch.cuda.synchronize(); dt=time.perf_counter()-t; print("H2D 512 x 4MB GB/s", n*4/dt/1e9)
