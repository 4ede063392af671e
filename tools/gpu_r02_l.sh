#!/bin/bash
# round 2, GPU call L: one warp per band chunk (k_populate_roles, bands 32 / 64) and 2-byte rows in the register traceback kernel
set -x
O=gpurun_out/r02l
mkdir -p $O
timeout 120 python tools/roles_smoke.py 64 > $O/smoke_roles.txt 2>&1; echo "rc=$?" >> $O/smoke_roles.txt
PHMM_NO_ROLE_WARPS=1 timeout 120 python tools/roles_smoke.py 64 > $O/smoke_lanes.txt 2>&1; echo "rc=$?" >> $O/smoke_lanes.txt
cat $O/smoke_roles.txt $O/smoke_lanes.txt
timeout 400 compute-sanitizer --tool racecheck --racecheck-report all python tools/roles_smoke.py 24 > $O/racecheck.txt 2>&1; echo "rc=$?" >> $O/racecheck.txt
tail -5 $O/racecheck.txt
timeout 900 python -m pytest tests/test_gpu_wide.py tests/test_gpu_fullsize.py tests/test_gpu_parity.py -m gpu -x -q > $O/pytest_subset.log 2>&1; echo "pytest rc=$?" >> $O/pytest_subset.log
tail -4 $O/pytest_subset.log | cut -c1-300
B="python bench.py --no-cpu-baseline"
timeout 300 $B --config C4 --steps 5 --warmup 3 > $O/bench_c4_roles.json 2> $O/bench_c4_roles.err
PHMM_NO_ROLE_WARPS=1 timeout 300 $B --config C4 --steps 5 --warmup 3 > $O/bench_c4_lanes.json 2> $O/bench_c4_lanes.err
timeout 400 $B --config C3 --steps 3 --warmup 3 --band 64 > $O/bench_c3_band64_roles.json 2> $O/bench_c3_band64_roles.err
timeout 300 python tools/bench_align.py 400000 > $O/align_fast.json 2> $O/align_fast.err
PHMM_NO_FAST_ALIGN=1 timeout 600 python tools/bench_align.py 400000 > $O/align_generic.json 2> $O/align_generic.err
cat $O/align_fast.json $O/align_generic.json
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
