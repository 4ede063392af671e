#!/bin/bash
# round 2, GPU call E: regions entry point, adapter e2e, headline capture (traffic), C1-shaped batches
set -x
O=gpurun_out/r02e
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -x -q -s > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -8 $O/pytest_gpu.log | cut -c1-400
B="python bench.py --no-cpu-baseline"
timeout 300 $B --config C1 --steps 20 --warmup 5 > $O/bench_c1_single.json 2> $O/bench_c1_single.err
timeout 300 $B --config C1 --batch-regions 1000 --steps 5 --warmup 3 > $O/bench_c1_x1000.json 2> $O/bench_c1_x1000.err
timeout 300 $B --config C1 --batch-regions 1000 --shortcut --map --flank 40,40 --steps 5 --warmup 3 > $O/bench_c1_x1000_prod.json 2> $O/bench_c1_x1000_prod.err
timeout 300 $B --config C1 --regions 200 --steps 5 --warmup 3 > $O/bench_c1_200calls.json 2> $O/bench_c1_200calls.err
timeout 300 $B --config C5 --regions 8 --steps 3 --warmup 2 > $O/bench_c5_8regions.json 2> $O/bench_c5.err
timeout 400 python bench.py --steps 3 --warmup 3 > $O/bench_default.json 2> $O/bench_default.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
NCU="ncu --set full --clock-control none --import-source on"
timeout 900 $NCU -k regex:k_populate_fast -s 1 -c 1 -o $O/fast16_c3 $B --config C3 --steps 1 --warmup 1 > $O/ncu_fast16.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_c3.csv $B --config C3 --steps 1 --warmup 1 > $O/launches_c3.log 2>&1
ls -la $O
