#!/bin/bash
# Where does the e2e leg's step-to-step variance come from?  Host topology of the box, then the file-bytes leg (config 2 shape,
# 256 MB chunks) twenty times unbound and twenty times bound to the CPUs next to GPU 0.  Run under gpurun; output in gpurun_out/.
out=gpurun_out/e2e_numa_probe.txt
mkdir -p gpurun_out
{
  echo "== lscpu"; lscpu | grep -i -E 'model name|socket|numa|^cpu\(s\)|thread'
  echo "== affinity"; python -c "import os; print(sorted(os.sched_getaffinity(0)))"
  echo "== cgroup cpu"; cat /sys/fs/cgroup/cpu.max 2>/dev/null
  echo "== topo"; nvidia-smi topo -m 2>&1 | head -20
  echo "== numa nodes"; ls /sys/devices/system/node 2>/dev/null | tr '\n' ' '; echo
  for n in /sys/devices/system/node/node*; do echo "$n: $(cat $n/cpulist) $(grep MemFree $n/meminfo)"; done
  echo "== unbound"; MODES=pp CHUNKS=256 REPS=20 python tools/e2e_probe.py
  echo "== bound"; BIND=1 MODES=pp CHUNKS=256 REPS=20 python tools/e2e_probe.py
  echo "== unbound again, chunks 128 / 512"; MODES=pp CHUNKS=128,512 REPS=12 python tools/e2e_probe.py
} > $out 2>&1
tail -30 $out
