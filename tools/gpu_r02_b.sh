#!/bin/bash
# round 2, GPU call B: the multi-lane band kernels (packed + 32-bit): parity tests, then bench lines and one ncu capture
set -x
O=gpurun_out/r02b
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -15 $O/pytest_gpu.log
B="python bench.py --no-cpu-baseline"
timeout 300 $B --config C3 --steps 3 --warmup 2 > $O/bench_c3.json 2> $O/bench_c3.err
timeout 300 $B --config C4 --steps 3 --warmup 2 > $O/bench_c4.json 2> $O/bench_c4.err
timeout 300 $B --config C2 --band 64 --hap-len 400 --steps 3 --warmup 2 > $O/bench_c2_band64.json 2> $O/bench_c2_band64.err
timeout 300 $B --config C2 --band 128 --hap-len 520 --reads 50000 --steps 3 --warmup 2 > $O/bench_c2_band128.json 2> $O/bench_c2_band128.err
timeout 300 $B --config C2 --band 256 --hap-len 800 --reads 20000 --steps 3 --warmup 2 > $O/bench_c2_band256.json 2> $O/bench_c2_band256.err
timeout 300 $B --config C2 --int-scores --steps 3 --warmup 2 > $O/bench_c2_int32.json 2> $O/bench_c2_int32.err
timeout 300 $B --config C2 --band 64 --int-scores --hap-len 400 --reads 50000 --steps 3 --warmup 2 > $O/bench_c2_band64_int32.json 2> $O/bench_c2_band64_int32.err
timeout 300 $B --config C2 --band 64 --read-lens 1000 --hap-len 1400 --reads 4000 --int-scores --steps 3 --warmup 2 > $O/bench_long1k_band64_int32.json 2> $O/bench_long1k_band64.err
timeout 300 $B --config C2 --band 128 --read-lens 5000 --hap-len 6000 --reads 1000 --haps 32 --int-scores --steps 2 --warmup 1 > $O/bench_long5k_band128_int32.json 2> $O/bench_long5k.err
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:k_populate_fast -s 1 -c 1 -o $O/fast32ml $B --config C4 --reads 100000 --steps 1 --warmup 1 > $O/ncu_fast32.log 2>&1
timeout 600 $NCU -k regex:k_populate_wide -s 1 -c 1 -o $O/wide64 $B --config C2 --band 64 --int-scores --hap-len 400 --reads 50000 --steps 1 --warmup 1 > $O/ncu_wide64.log 2>&1
ls -la $O
