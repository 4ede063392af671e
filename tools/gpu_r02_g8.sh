#!/bin/bash
# round 2, 8-GPU call: weak scaling (C3 per rank), strong scaling (C3's 1M reads split over the ranks, matrix re-assembled on rank 0),
# C5 as written (1k regions over 8 GPUs = 125 regions per rank, per-region gather overlapped with the next region's compute)
set -x
O=gpurun_out/r02g
mkdir -p $O
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > $O/smi.txt
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
B="bench.py --no-cpu-baseline --steps 5 --warmup 3"
timeout 400 python $B --config C3 > $O/c3_n1.json 2> $O/c3_n1.err
timeout 400 $TR --nproc-per-node 8 --master-port 29511 $B --gpus 8 --config C3 > $O/c3_weak_n8.json 2> $O/c3_weak_n8.err
timeout 400 $TR --nproc-per-node 8 --master-port 29512 $B --gpus 8 --config C3 --scaling strong > $O/c3_strong_n8.json 2> $O/c3_strong_n8.err
timeout 400 $TR --nproc-per-node 4 --master-port 29513 $B --gpus 4 --config C3 --scaling strong > $O/c3_strong_n4.json 2> $O/c3_strong_n4.err
timeout 400 $TR --nproc-per-node 8 --master-port 29514 bench.py --no-cpu-baseline --steps 3 --warmup 2 --gpus 8 --config C5 > $O/c5_n8.json 2> $O/c5_n8.err
timeout 400 python bench.py --no-cpu-baseline --steps 3 --warmup 2 --config C5 > $O/c5_n1.json 2> $O/c5_n1.err
tail -c 600 $O/*.err
ls -la $O
