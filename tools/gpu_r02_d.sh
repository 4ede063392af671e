#!/bin/bash
# round 2, GPU call D: mapper v3 + lean flank kernel with FMA-pipe adds: parity tests, production-mode bench lines
set -x
O=gpurun_out/r02d
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -x -q -s > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -8 $O/pytest_gpu.log
B="python bench.py --no-cpu-baseline"
timeout 300 $B --config C2 --flank 60,60 --steps 3 --warmup 2 > $O/bench_c2_flank.json 2> $O/bench_c2_flank.err
timeout 300 $B --config C2 --shortcut --map --steps 3 --warmup 2 > $O/bench_c2_refmode.json 2> $O/bench_c2_refmode.err
timeout 300 $B --config C2 --shortcut --map --flank 60,60 --steps 3 --warmup 2 > $O/bench_c2_prod.json 2> $O/bench_c2_prod.err
timeout 300 $B --config C2 --shortcut --map --flank 60,60 --error-model PCR-free.HiSeq-2500 --steps 3 --warmup 2 > $O/bench_c2_prod_errmodel.json 2> $O/bench_c2_prod_errmodel.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_c2_prod.csv $B --config C2 --shortcut --map --flank 60,60 --steps 1 --warmup 1 > $O/launches_prod.log 2>&1
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:k_kmer_map -s 0 -c 1 -o $O/kmermap $B --config C2 --shortcut --map --steps 1 --warmup 1 > $O/ncu_kmer.log 2>&1
timeout 600 $NCU -k regex:k_populate_flank_acc -s 1 -c 1 -o $O/flankacc16 $B --config C2 --flank 60,60 --steps 1 --warmup 1 > $O/ncu_flankacc.log 2>&1
ls -la $O
