#!/bin/bash
# round-2 GPU call 24 (4 GPUs): scaling at HEAD, N = 4
cd $GRAFT_REPO_ROOT
S=gpurun_out/r24_status.txt; : > $S
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 4 --steps 2000 --warmup 20 > gpurun_out/r24_bench_4gpu.json 2> gpurun_out/r24_bench_4gpu.err; echo "N=4 rc=$?" >> $S
