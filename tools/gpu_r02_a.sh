#!/bin/bash
# round 2, GPU call A: sanity (pytest -m gpu) + ncu captures of the kernels round 1 never profiled, on the round-1 code
set -x
mkdir -p gpurun_out/r02a
O=gpurun_out/r02a
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
B="python bench.py --no-cpu-baseline"
# plain bench lines first (never under ncu)
timeout 300 $B --config C2 --flank 60,60 --steps 3 --warmup 2 > $O/bench_c2_flank.json 2> $O/bench_c2_flank.err
timeout 300 $B --config C2 --shortcut --map --steps 3 --warmup 2 > $O/bench_c2_refmode.json 2> $O/bench_c2_refmode.err
timeout 300 $B --config C2 --shortcut --map --flank 60,60 --steps 3 --warmup 2 > $O/bench_c2_prod.json 2> $O/bench_c2_prod.err
timeout 300 $B --config C4 --steps 3 --warmup 2 > $O/bench_c4.json 2> $O/bench_c4.err
timeout 300 $B --config C2 --band 64 --reads 20000 --steps 2 --warmup 1 > $O/bench_c2_band64.json 2> $O/bench_c2_band64.err
timeout 300 $B --config C2 --int-scores --reads 20000 --steps 2 --warmup 1 > $O/bench_c2_int32.json 2> $O/bench_c2_int32.err
timeout 300 $B --config C2 --band 64 --read-lens 1000 --hap-len 1400 --reads 2000 --int-scores --steps 2 --warmup 1 > $O/bench_long_band64.json 2> $O/bench_long_band64.err
# ncu --set full, one launch each
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:k_populate_flank -s 1 -c 1 -o $O/flank16 $B --config C2 --flank 60,60 --steps 1 --warmup 1 > $O/ncu_flank.log 2>&1
timeout 600 $NCU -k regex:k_kmer_map -s 1 -c 1 -o $O/kmermap $B --config C2 --shortcut --map --steps 1 --warmup 1 > $O/ncu_kmer.log 2>&1
timeout 600 $NCU -k regex:k_populate_fast -s 1 -c 1 -o $O/fast32 $B --config C4 --reads 100000 --steps 1 --warmup 1 > $O/ncu_fast32.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_c2_prod.csv $B --config C2 --shortcut --map --flank 60,60 --steps 1 --warmup 1 > $O/launches_prod.log 2>&1
ls -la $O
