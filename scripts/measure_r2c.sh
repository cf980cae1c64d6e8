set -x
timeout 400 python -m pytest tests/test_gpu_layers.py tests/test_gpu_configs.py -x -q 2>&1 | tail -6
timeout 100 python scripts/profile_step.py > gpurun_out/step_profile_lean3.json 2>/dev/null
timeout 200 python bench.py --steps 100 --warmup 5 --no-tc-rooflines --no-cpu-baseline > gpurun_out/bench_lean3.json; cut -c1-200 gpurun_out/bench_lean3.json
timeout 100 python scripts/ncu_step.py && timeout 400 ncu --set full --clock-control none --import-source on -k 'regex:k_gyro|k_mobius|k_mob_|k_hradius|k_absmax|k_split2h|k_bce|k_colsum' --launch-skip 86 --launch-count 30 -f -o gpurun_out/r2c_step_kernels python scripts/ncu_step.py > gpurun_out/ncu_step.log 2>&1
tail -2 gpurun_out/ncu_step.log
