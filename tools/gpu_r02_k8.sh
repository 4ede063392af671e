#!/bin/bash
# round 2, second 8-GPU call: peer-store output (every rank's epilogue writes into rank 0's HBM, one barrier per step) —
# weak C3, strong C3, C5 (125 regions per rank); single-GPU lines of the same build beside them
set -x
O=gpurun_out/r02k
mkdir -p $O
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > $O/smi.txt
nvidia-smi topo -m > $O/topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
B="bench.py --no-cpu-baseline --steps 5 --warmup 3"
timeout 300 python $B --config C3 > $O/c3_n1.json 2> $O/c3_n1.err
timeout 300 $TR --nproc-per-node 8 --master-port 29811 $B --gpus 8 --config C3 --scaling strong > $O/c3_strong_n8.json 2> $O/c3_strong_n8.err
timeout 300 $TR --nproc-per-node 8 --master-port 29812 $B --gpus 8 --config C3 > $O/c3_weak_n8.json 2> $O/c3_weak_n8.err
timeout 400 $TR --nproc-per-node 8 --master-port 29814 bench.py --no-cpu-baseline --steps 3 --warmup 2 --gpus 8 --config C5 > $O/c5_n8.json 2> $O/c5_n8.err
timeout 300 python bench.py --no-cpu-baseline --steps 2 --warmup 1 --config C5 > $O/c5_n1.json 2> $O/c5_n1.err
tail -c 400 $O/*.err
for f in $O/*.json; do python - "$f" <<'PY'
import sys, json
f=sys.argv[1]
try:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, d['scaling'], 'value %.0f ms/step %.2f kernel %.2f barrier/region %.2f parity %s check %s clocks %s' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['gather_ms_per_region'], d['parity']['mismatches'], d.get('gather_check',{}).get('ranks_mismatched'), d['clocks'].get('sm_mhz')))
except Exception as e:
    print(f, 'ERR', e)
PY
done
