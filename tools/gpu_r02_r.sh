#!/bin/bash
# round 2, GPU call R (final tree): the whole GPU suite, smoke(), the default bench line, non-headline lines of the same build,
# launch list of a flank-state step, compute-sanitizer memcheck over the forward / backward flank kernels
set -x
O=gpurun_out/r02r
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -4 $O/pytest_gpu.log | cut -c1-300
timeout 120 python __graft_entry__.py --smoke > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log; tail -2 $O/smoke.log
timeout 300 python bench.py > $O/bench_default.json 2> $O/bench_default.err
B="python bench.py --no-cpu-baseline --steps 3 --warmup 3"
timeout 120 $B --config C4 > $O/bench_c4.json 2> $O/bench_c4.err
timeout 120 $B --config C2 --flank 60,60 > $O/bench_c2_flank.json 2> $O/bench_c2_flank.err
timeout 120 $B --config C2 --shortcut --map --flank 60,60 > $O/bench_c2_prod.json 2> $O/bench_c2_prod.err
timeout 120 $B --config C3 --flank 60,60 > $O/bench_c3_flank.json 2> $O/bench_c3_flank.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_c2_flank.csv $B --config C2 --flank 60,60 --steps 1 --warmup 1 > $O/launches_flank.log 2>&1
timeout 100 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_flank_fb.py -m gpu -x -q -k "16 and regions" > $O/memcheck_flank.log 2>&1; echo "memcheck rc=$?" >> $O/memcheck_flank.log
tail -5 $O/memcheck_flank.log | cut -c1-300
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
ls -la $O
