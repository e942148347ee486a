#!/bin/bash
# Round-end measurement on the GPU box (run under gpurun): tests, bench (both arms), per-kernel rooflines, ncu launch list and
# one --set full capture of each dominant kernel.  Everything lands in gpurun_out/<tag>_*.  Usage: tools/profile_round.sh <tag>
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
(time timeout 900 python -m pytest tests -m gpu -x -q) > $out/${tag}_tests.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1
timeout 900 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err
timeout 900 python bench.py --impl reference > $out/${tag}_bench_reference.json 2>> $out/${tag}_bench.err
timeout 900 python tools/kernel_rooflines.py > $out/${tag}_kernel_rooflines.json 2> $out/${tag}_kernel_rooflines.err
NCU="ncu --clock-control none"
timeout 900 $NCU --metrics gpu__time_duration.sum -k regex:'f9|umma|hankel|short|tail|poly|peak|stats' -c 400 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > $out/${tag}_ncu_launch.log 2>&1
timeout 900 $NCU --set full --import-source on -k regex:umma_fir -s 4 -c 1 -f -o $out/prof_${tag}_umma \
    python bench.py --steps 2 --warmup 3 --kernel-only > $out/${tag}_ncu_umma.log 2>&1
timeout 600 $NCU --set full --import-source on -k regex:hankel_fir -s 2 -c 1 -f -o $out/prof_${tag}_hankel \
    python tools/rate_bench.py 0 48000:192000 > $out/${tag}_ncu_hankel.log 2>&1
timeout 600 $NCU --set full --import-source on -k regex:short_kernel -s 2 -c 1 -f -o $out/prof_${tag}_short \
    python tools/rate_bench.py 1 44100:48000 > $out/${tag}_ncu_short.log 2>&1
tail -3 $out/${tag}_tests.log; cat $out/${tag}_smoke.log | tail -1; ls -la $out | grep ${tag}
