set -x
timeout 200 python scripts/profile_step.py > gpurun_out/step_profile.json 2> gpurun_out/step_profile.err
timeout 120 python scripts/ncu_x2.py && timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_x2_gemm --launch-skip 5 --launch-count 5 -f -o gpurun_out/r2_x2 python scripts/ncu_x2.py > gpurun_out/ncu_x2.log 2>&1
tail -2 gpurun_out/ncu_x2.log
timeout 200 python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-tc-rooflines > gpurun_out/plain.log 2>&1 && timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 600 --csv --log-file gpurun_out/launches_r2_cfg2.csv python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-tc-rooflines > gpurun_out/ncu_launch.log 2>&1
tail -1 gpurun_out/ncu_launch.log | cut -c1-200
timeout 300 python bench.py --workload cfg5 --steps 10 --warmup 3 > gpurun_out/bench_cfg5.json 2> gpurun_out/bench_cfg5.err; cut -c1-600 gpurun_out/bench_cfg5.json
