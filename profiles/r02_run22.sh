#!/bin/bash
# round-2 GPU call 22 (8 GPUs): scaling at HEAD -- weak-scaled cfg2 headline (streaming protocol) + config 5 strong-scaled over 8 ranks
cd $GRAFT_REPO_ROOT
S=gpurun_out/r22_status.txt; : > $S
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 2000 --warmup 20 > gpurun_out/r22_bench_8gpu.json 2> gpurun_out/r22_bench_8gpu.err; echo "N=8 rc=$?" >> $S
