#!/bin/bash
# round 2, 2-GPU call: peer-store output (CUDA IPC) — tests, then strong / weak / C5-shaped runs, peer vs nccl
set -x
O=gpurun_out/r02i
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_peer.py tests/test_gpu_wide.py -q -m gpu -k "peer or window or processes or reserved" > $O/pytest.log 2>&1
tail -n 5 $O/pytest.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 2"
B="bench.py --no-cpu-baseline --steps 4 --warmup 3 --gpus 2"
P=29700
for g in peer nccl; do
  P=$((P+1)); timeout 300 $TR --master-port $P $B --config C3 --scaling strong --gather $g > $O/c3_strong_n2_$g.json 2> $O/c3_strong_n2_$g.err
  P=$((P+1)); timeout 300 $TR --master-port $P $B --config C3 --gather $g > $O/c3_weak_n2_$g.json 2> $O/c3_weak_n2_$g.err
  P=$((P+1)); timeout 300 $TR --master-port $P bench.py --no-cpu-baseline --steps 2 --warmup 1 --gpus 2 --config C5 --regions 40 --gather $g > $O/c5_n2_$g.json 2> $O/c5_n2_$g.err
done
tail -c 600 $O/*peer.err
for f in $O/*.json; do python - "$f" <<'PY'
import sys, json
f=sys.argv[1]
try:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, d['scaling'], 'value %.0f ms/step %.2f kernel %.2f gather/region %.2f parity %s check %s' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['gather_ms_per_region'], d['parity']['mismatches'], d.get('gather_check')))
except Exception as e:
    print(f, 'ERR', e)
PY
done
