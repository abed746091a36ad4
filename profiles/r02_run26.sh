#!/bin/bash
# round-2 GPU call 26: the default bench line and the driver-like line at HEAD (compact e2e format, streaming region behind a flush, new traffic.json), smoke, stream + host-step tests
cd $GRAFT_REPO_ROOT
S=gpurun_out/r26_status.txt; : > $S
timeout 900 python -m pytest tests/test_cuda_streams.py tests/test_cuda_oracle.py -m gpu -q -x -k "stream or pdl or pipelined or sliced or fixed_stride" > gpurun_out/r26_tests.log 2>&1; echo "tests rc=$?" >> $S
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r26_smoke.log 2>&1; echo "smoke rc=$?" >> $S
timeout 1500 python bench.py > gpurun_out/r26_bench_default.json 2> gpurun_out/r26_bench_default.err; echo "bench default rc=$?" >> $S
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r26_bench_driverlike.json 2> gpurun_out/r26_bench_driverlike.err; echo "bench driver-like rc=$?" >> $S
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r26_bench_reference.json 2> gpurun_out/r26_bench_reference.err; echo "bench reference rc=$?" >> $S
