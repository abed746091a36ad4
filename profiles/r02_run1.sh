#!/bin/bash
# round-2 GPU call 1: tests, reset costs, headline bench variants
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r1_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --maxfail=25 -x -k "reset_data or generate" > gpurun_out/r1_tests_a.log 2>&1
echo "tests_a rc=$?" >> gpurun_out/r1_status.txt
timeout 300 python profiles/reset_costs.py 2048 > gpurun_out/r1_reset_costs.jsonl 2> gpurun_out/r1_reset_costs.err
echo "reset_costs rc=$?" >> gpurun_out/r1_status.txt
timeout 300 python bench.py --only-headline --steps 256 --no-cpu > gpurun_out/r1_bench_pipe.json 2> gpurun_out/r1_bench_pipe.err
echo "bench pipelined rc=$?" >> gpurun_out/r1_status.txt
timeout 300 python bench.py --only-headline --steps 256 --no-cpu --e2e single > gpurun_out/r1_bench_single.json 2> gpurun_out/r1_bench_single.err
echo "bench single rc=$?" >> gpurun_out/r1_status.txt
GE_LANE_T=32 timeout 300 python bench.py --only-headline --steps 256 --no-cpu > gpurun_out/r1_bench_T32.json 2> gpurun_out/r1_bench_T32.err
GE_LANE_T=128 timeout 300 python bench.py --only-headline --steps 256 --no-cpu > gpurun_out/r1_bench_T128.json 2> gpurun_out/r1_bench_T128.err
timeout 1800 python -m pytest tests -m gpu -q --maxfail=25 -k "not reset_data and not generate" > gpurun_out/r1_tests_b.log 2>&1
echo "tests_b rc=$?" >> gpurun_out/r1_status.txt
