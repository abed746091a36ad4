#!/bin/bash
# round-2 GPU call 13 (8 GPUs): the scaling curve -- same command at N = 8 (ranks pinned to core blocks / unpinned), 4, 2
cd $GRAFT_REPO_ROOT
S=gpurun_out/r13_status.txt; : > $S
run() {  # N port extra-args tag
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus $1 --steps 200 --warmup 5 $3 > gpurun_out/r13_bench_$4.json 2> gpurun_out/r13_bench_$4.err; echo "N=$1 $4 rc=$?" >> $S
}
run 8 29521 "" 8gpu
run 8 29522 "--no-pin" 8gpu_nopin
run 4 29523 "" 4gpu
run 2 29524 "" 2gpu
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29525 bench.py --impl reference --gpus 8 --steps 20 --warmup 5 > gpurun_out/r13_bench_reference_8.json 2> gpurun_out/r13_bench_reference_8.err; echo "reference N=8 rc=$?" >> $S
