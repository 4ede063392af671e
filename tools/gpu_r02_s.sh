#!/bin/bash
# round 2, GPU call S: lane groups in the labelled flank kernel (the packed kernels' ties: one or two candidates per read) — flank /
# parity / wide suites and the flank-state bench lines
set -x
O=gpurun_out/r02s
mkdir -p $O
timeout 200 python -m pytest tests/test_gpu_flank_fb.py tests/test_gpu_parity.py tests/test_gpu_wide.py -m gpu -x -q > $O/pytest_gpu_flank.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_flank.log
tail -4 $O/pytest_gpu_flank.log | cut -c1-300
B="python bench.py --no-cpu-baseline --steps 3 --warmup 3"
timeout 60 $B --config C2 --flank 60,60 > $O/bench_c2_flank.json 2> $O/bench_c2_flank.err
timeout 60 $B --config C2 --shortcut --map --flank 60,60 > $O/bench_c2_prod.json 2> $O/bench_c2_prod.err
timeout 60 $B --config C3 --flank 60,60 > $O/bench_c3_flank.json 2> $O/bench_c3_flank.err
timeout 60 $B --config C1 --batch-regions 1000 --flank 60,60 > $O/bench_c1x1000_flank.json 2> $O/bench_c1x1000_flank.err
for f in $O/bench_*.json; do python - "$f" <<'PY'
import sys, json
f=sys.argv[1]
try:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, 'value %.0f e2e %.0f ms/step %.2f kernel %.2f parity %s launches %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['parity'].get('mismatches'), d.get('gpu_launches')))
except Exception as e:
    print(f, 'ERR', e)
PY
done
