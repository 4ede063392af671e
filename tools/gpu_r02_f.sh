#!/bin/bash
# round 2, GPU call F: register traceback + N4 tests, align throughput before/after, headline re-check after the store-path fix
set -x
O=gpurun_out/r02f
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -x -q -s > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -8 $O/pytest_gpu.log | cut -c1-400
timeout 300 python tools/bench_align.py 400000 > $O/align_fast.json 2> $O/align_fast.err
PHMM_NO_FAST_ALIGN=1 timeout 600 python tools/bench_align.py 400000 > $O/align_generic.json 2> $O/align_generic.err
B="python bench.py --no-cpu-baseline"
timeout 300 $B --config C3 --steps 5 --warmup 3 > $O/bench_c3.json 2> $O/bench_c3.err
timeout 300 $B --config C2 --steps 5 --warmup 3 > $O/bench_c2.json 2> $O/bench_c2.err
NCU="ncu --set full --clock-control none --import-source on"
timeout 900 $NCU -k regex:k_populate_fast -s 1 -c 1 -o $O/fast16_c3 $B --config C3 --steps 1 --warmup 1 > $O/ncu_fast16.log 2>&1
timeout 600 $NCU -k regex:k_align_reads_fast -s 1 -c 1 -o $O/alignfast python tools/bench_align.py 100000 > $O/ncu_align.log 2>&1
cat $O/align_fast.json $O/align_generic.json
ls -la $O
