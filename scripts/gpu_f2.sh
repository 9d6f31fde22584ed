mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q -k "prover or standins" > gpurun_out/f2_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/f2_gputests.log; tail -4 gpurun_out/f2_gputests.log
timeout 200 python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/f2_bench.log 2> gpurun_out/f2_bench.err; echo "bench rc=$?"; tail -1 gpurun_out/f2_bench.log | cut -c1-300
timeout 150 ncu --set full --clock-control none -k regex:expr_kernel -s 55 -c 1 -o gpurun_out/f2_expr -f python scripts/proof_trace.py super 20 128 > gpurun_out/f2_ncu_expr.log 2>&1
ls -la gpurun_out | tail -5
