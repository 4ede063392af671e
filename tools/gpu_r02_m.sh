#!/bin/bash
# round 2, GPU call M: whole GPU suite on the current build, then the occupancy variants of the packed kernels
# (launch bounds: resident blocks per SM for band 16 / band 8 / bands >= 32) on C3, C4, C1 x 1000 regions and C2
set -x
O=gpurun_out/r02m
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -5 $O/pytest_gpu.log | cut -c1-300
B="python bench.py --no-cpu-baseline --steps 5 --warmup 3"
cp octopus_b200/libphmm_b200.so /tmp/lib_default.so
for v in 444 564 664 565; do
  cp tools/_variants/lib_mb$v.so octopus_b200/libphmm_b200.so
  timeout 200 $B --config C3 > $O/v${v}_c3.json 2> $O/v${v}_c3.err
  timeout 200 $B --config C4 > $O/v${v}_c4.json 2> $O/v${v}_c4.err
  timeout 200 $B --config C1 --batch-regions 1000 > $O/v${v}_c1x1000.json 2> $O/v${v}_c1x1000.err
  timeout 200 $B --config C2 > $O/v${v}_c2.json 2> $O/v${v}_c2.err
done
cp /tmp/lib_default.so octopus_b200/libphmm_b200.so
for f in $O/v*.json; do python - "$f" <<'PY'
import sys, json
f=sys.argv[1]
try:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, 'value %.0f e2e %.0f ms/step %.2f kernel %.2f kernel_gcups %.0f parity %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['kernel_gcups'], d['parity']['mismatches']))
except Exception as e:
    print(f, 'ERR', e)
PY
done
