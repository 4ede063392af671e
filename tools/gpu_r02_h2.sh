#!/bin/bash
# round 2, 2-GPU call: does keeping SMs free of persistent DP blocks (phmm_reserve_sms) let the NCCL gather run beside the next populate?
# strong scaling on C3 and per-region gathers on a C5-shaped job, reservation 0 / 8 / 16; plus the reserved-SMs GPU test.
set -x
O=gpurun_out/r02h
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 2"
B="bench.py --no-cpu-baseline --steps 4 --warmup 3 --gpus 2"
timeout 300 python -m pytest tests/test_gpu_wide.py -q -m gpu -k "reserved_sms" > $O/pytest.log 2>&1
P=29600
for rs in 0 8 16; do
  P=$((P+1)); timeout 300 $TR --master-port $P $B --config C3 --scaling strong --reserve-sms $rs > $O/c3_strong_n2_rs$rs.json 2> $O/c3_strong_n2_rs$rs.err
  P=$((P+1)); timeout 300 $TR --master-port $P $B --config C3 --reserve-sms $rs > $O/c3_weak_n2_rs$rs.json 2> $O/c3_weak_n2_rs$rs.err
  P=$((P+1)); timeout 300 $TR --master-port $P bench.py --no-cpu-baseline --steps 2 --warmup 1 --gpus 2 --config C5 --regions 40 --reserve-sms $rs > $O/c5_n2_rs$rs.json 2> $O/c5_n2_rs$rs.err
done
tail -n 3 $O/pytest.log
tail -c 300 $O/*.err
cat $O/*.json | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    print(d['config']['workload'][:40], d['scaling'], 'rs', d['config']['reserved_sms'], 'value %.0f ms/step %.2f kernel %.2f gather/region %.2f parity %s' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['gather_ms_per_region'], d['parity']['mismatches']))
"
