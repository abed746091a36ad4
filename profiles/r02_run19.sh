#!/bin/bash
# round-2 GPU call 19: ncu capture of the DistributionCenter step kernel with fixed-stride rows
cd $GRAFT_REPO_ROOT
NCU="ncu --set full --clock-control none --import-source on"
timeout 900 $NCU -k regex:dc_step --launch-skip 40 -c 2 -f -o gpurun_out/r19_ncu_cfg5_distcenter python bench.py --workload cfg5_distcenter --only-headline --no-streaming --steps 64 --warmup 3 --no-cpu --e2e-steps 3 --no-e2e-obs > gpurun_out/r19_ncu.log 2>&1; echo "rc=$?" > gpurun_out/r19_status.txt
