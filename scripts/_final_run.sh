set -x
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 400 python bench.py --steps 100 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 400 gpurun_out/bench_n1.err
timeout 300 python scripts/microbench_kernels.py --logB 18 > gpurun_out/microbench.json 2> gpurun_out/microbench.err
timeout 200 python scripts/profile_step.py > gpurun_out/step_profile.json 2> gpurun_out/step_profile.err
timeout 100 python scripts/ncu_x3.py > gpurun_out/ncu_x3_plain.log 2>&1 && timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_tc_gemm --launch-skip 5 --launch-count 5 -o gpurun_out/r1_x3_gemm -f python scripts/ncu_x3.py > gpurun_out/ncu_x3.log 2>&1
tail -2 gpurun_out/ncu_x3.log
