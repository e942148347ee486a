#!/bin/bash
# Sweep banded-kernel configurations (TK,TA,nGB) on the default workload; prints one JSON line per config.
for cfg in "$@"; do
  F9_BANDED_CFG="$cfg" python bench.py --steps 3 --warmup 3 --files 64 --kernel-only 2>&1 | tail -1
done
