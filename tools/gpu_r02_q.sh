#!/bin/bash
# round 2, GPU call Q: flank kernels with lane groups (regions with few haplotypes), the backward pass without an end-row body
# (zeros below row L) and the packed two-sweep crossing-cell decision: parity tests, bench lines, ncu of both kernels
set -x
O=gpurun_out/r02q
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_flank_fb.py tests/test_gpu_parity.py tests/test_gpu_wide.py -m gpu -x -q > $O/pytest_gpu_flank.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_flank.log
tail -6 $O/pytest_gpu_flank.log | cut -c1-400
B="python bench.py --no-cpu-baseline --steps 3 --warmup 3"
timeout 200 $B --config C2 --flank 60,60 > $O/bench_c2_flank.json 2> $O/bench_c2_flank.err
timeout 200 $B --config C2 --shortcut --map --flank 60,60 > $O/bench_c2_prod.json 2> $O/bench_c2_prod.err
timeout 200 $B --config C3 --flank 60,60 > $O/bench_c3_flank.json 2> $O/bench_c3_flank.err
timeout 200 $B --config C1 --batch-regions 1000 --flank 60,60 > $O/bench_c1x1000_flank.json 2> $O/bench_c1x1000_flank.err
timeout 200 $B --config C1 --batch-regions 1000 --flank 60,60 --shortcut --map > $O/bench_c1x1000_prod.json 2> $O/bench_c1x1000_prod.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_flank_ -s 2 -c 2 -o $O/flank_fwd_bwd16 $B --config C2 --flank 60,60 --steps 1 --warmup 1 > $O/ncu_flankfb.log 2>&1
for f in $O/bench_*.json; do python - "$f" <<'PY'
import sys, json
f=sys.argv[1]
try:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, 'value %.0f e2e %.0f ms/step %.2f kernel %.2f parity %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['parity'].get('mismatches')))
except Exception as e:
    print(f, 'ERR', e)
PY
done
ls -la $O
