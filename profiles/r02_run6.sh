#!/bin/bash
# round-2 GPU call 6: DistributionCenter transposed mask (tests + bench), e2e policy variants
cd $GRAFT_REPO_ROOT
S=gpurun_out/r6_status.txt; : > $S
timeout 1200 python -m pytest tests -m gpu -q --maxfail=20 -k "Distribution or cfg5 or golden or pool or properties" > gpurun_out/r6_tests.log 2>&1; echo "tests rc=$?" >> $S
timeout 600 python bench.py --workload cfg5_distcenter --only-headline --steps 100 --no-cpu --e2e-steps 10 --no-e2e-obs > gpurun_out/r6_bench_dc.json 2> gpurun_out/r6_bench_dc.err; echo "bench dc rc=$?" >> $S
timeout 300 python bench.py --only-headline --steps 256 --no-cpu --no-e2e-obs > gpurun_out/r6_cfg2_hostpolicy.json 2> gpurun_out/r6_cfg2.err
timeout 300 python bench.py --only-headline --steps 256 --no-cpu --no-e2e-obs --e2e-device-policy > gpurun_out/r6_cfg2_devpolicy.json 2>> gpurun_out/r6_cfg2.err
timeout 300 python bench.py --only-headline --steps 256 --no-cpu --no-e2e-obs --e2e-device-policy --e2e-chunks 3 > gpurun_out/r6_cfg2_devpolicy_c3.json 2>> gpurun_out/r6_cfg2.err
timeout 300 python bench.py --only-headline --steps 256 --no-cpu --no-e2e-obs --e2e single > gpurun_out/r6_cfg2_single.json 2>> gpurun_out/r6_cfg2.err
echo "cfg2 variants done" >> $S
NCU="ncu --set full --clock-control none --import-source on"
timeout 900 $NCU -k regex:dc_step -c 2 -f -o gpurun_out/r6_ncu_dc python bench.py --workload cfg5_distcenter --only-headline --steps 4 --warmup 3 --no-cpu --e2e-steps 3 --no-e2e-obs > gpurun_out/r6_ncu_dc.log 2>&1; echo "ncu dc rc=$?" >> $S
