#!/bin/bash
# round 2, GPU call J: the shorter deletion update (OGE: gap_open >= gap_extend) + additions on the FMA pipe in the packed kernels;
# whole GPU test suite, headline with ordered / unordered penalties, C4, C1 regions batch, production C2, ncu of the headline kernel
set -x
O=gpurun_out/r02j
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -6 $O/pytest_gpu.log | cut -c1-300
B="python bench.py --no-cpu-baseline"
timeout 300 $B --config C3 --steps 5 --warmup 3 > $O/bench_c3.json 2> $O/bench_c3.err
timeout 300 $B --config C3 --steps 5 --warmup 3 --unordered-penalties > $O/bench_c3_unordered.json 2> $O/bench_c3_unordered.err
timeout 300 $B --config C4 --steps 5 --warmup 3 > $O/bench_c4.json 2> $O/bench_c4.err
timeout 300 $B --config C4 --steps 5 --warmup 3 --unordered-penalties > $O/bench_c4_unordered.json 2> $O/bench_c4_unordered.err
timeout 300 $B --config C1 --batch-regions 1000 --steps 5 --warmup 3 > $O/bench_c1x1000.json 2> $O/bench_c1x1000.err
timeout 300 $B --config C2 --steps 5 --warmup 3 --shortcut --map --flank 30,40 > $O/bench_c2_prod.json 2> $O/bench_c2_prod.err
timeout 300 $B --config C3 --steps 5 --warmup 3 --band 64 > $O/bench_c3_band64.json 2> $O/bench_c3_band64.err
timeout 300 $B --config C3 --steps 5 --warmup 3 --int-scores > $O/bench_c3_int.json 2> $O/bench_c3_int.err
NCU="ncu --set full --clock-control none --import-source on"
timeout 900 $NCU -k regex:k_populate_fast -s 1 -c 1 -o $O/fast16_c3 $B --config C3 --steps 1 --warmup 1 > $O/ncu_fast16.log 2>&1
timeout 900 $NCU -k regex:k_populate_fast -s 1 -c 1 -o $O/fast32_c4 $B --config C4 --steps 1 --warmup 1 > $O/ncu_fast32.log 2>&1
for f in $O/bench_*.json; do python - "$f" <<'PY'
import sys, json
f=sys.argv[1]
try:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, 'value %.0f e2e %.0f ms/step %.2f kernel %.2f kernel_gcups %.0f parity %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['kernel_gcups'], d['parity']['mismatches']))
except Exception as e:
    print(f, 'ERR', e)
PY
done
ls -la $O
