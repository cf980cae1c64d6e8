set -x
timeout 500 python -m pytest tests/test_gpu_hradius.py tests/test_gpu_layers.py tests/test_gpu_noise.py tests/test_gpu_optim.py "tests/test_gpu_configs.py::test_cfg2_step_full_size" -x -q 2>&1 | tail -6
timeout 100 python scripts/profile_step.py > gpurun_out/step_profile_r2d.json 2>/dev/null
timeout 200 python bench.py --steps 100 --warmup 5 --no-tc-rooflines --no-cpu-baseline > gpurun_out/bench_r2d.json; cut -c1-200 gpurun_out/bench_r2d.json
