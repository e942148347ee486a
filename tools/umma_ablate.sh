#!/bin/bash
# Ablations of umma_fir_kernel on a -DF9_DIAG build (make -C f9-juce-resampler-studio_b200 DIAG=1), 96 -> 44.1 kHz, 512 channels of 10 s.
# F9_UMMA_DBG bits: 1 no MMAs, 2 no copies (register loader), 4 no stores, 8 no L2 prefetch, 16 prefetch two tiles ahead, 32 every tile
# re-reads the CTA's first tile (boxes always hit L2).  Results are wrong by construction; only the times mean something.
export F9DSP_DIAG_LIB=1
for d in ${DBGS:-0 8 16 32 5 13 37 36 33}; do
  echo -n "dbg=$d: "; F9_UMMA_DBG=$d python tools/rate_bench.py 0 96000:44100 2>&1 | tail -1
done
