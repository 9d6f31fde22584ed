#!/bin/bash
# usage: scripts/gpu_job.sh TAG  -- full GPU test suite, short bench, pipes microbenchmark and ncu captures into gpurun_out/TAG_*
T=$1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${T}_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_gputests.log; tail -8 gpurun_out/${T}_gputests.log
timeout 300 python bench.py --steps 20 --warmup 3 ${BENCH_FLAGS:---no-proof} > gpurun_out/${T}_bench.log 2>&1; tail -1 gpurun_out/${T}_bench.log | cut -c1-1200
if [ -x scripts/microbench/pipes ] && [ -n "$PIPES" ]; then timeout 60 ./scripts/microbench/pipes > gpurun_out/${T}_pipes.txt 2>&1; fi
if [ -n "$NCU_NTT" ]; then timeout 300 ncu --set full --clock-control none --import-source on -k regex:ntt_tile_kernel -s 3 -c 3 -o gpurun_out/${T}_ntt -f python scripts/prof_kernels.py ntt > gpurun_out/${T}_ncu_ntt.log 2>&1; fi
if [ -n "$NCU_MSM" ]; then timeout 300 ncu --set full --clock-control none --import-source on -k regex:"msm_acc_chunk|msm_wsum_level|msm_acc_levelN|msm_digits" -s 8 -c 8 -o gpurun_out/${T}_msm -f python scripts/prof_kernels.py msm > gpurun_out/${T}_ncu_msm.log 2>&1; fi
ls -la gpurun_out | tail -12
