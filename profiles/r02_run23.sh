#!/bin/bash
# round-2 GPU call 23 (2 GPUs): scaling at HEAD, N = 2
cd $GRAFT_REPO_ROOT
S=gpurun_out/r23_status.txt; : > $S
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 2000 --warmup 20 > gpurun_out/r23_bench_2gpu.json 2> gpurun_out/r23_bench_2gpu.err; echo "N=2 rc=$?" >> $S
