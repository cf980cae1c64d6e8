# Round-end evidence on one GPU (run under gpurun; outputs in gpurun_out/, the keepers are copied to profiles/):
#   1. the full GPU test suite (+ the strict parity audit it writes) and smoke()
#   2. the bench line of the metric's config with CPU baseline and tc_rooflines; the in-situ step profile
#   3. ncu launch list of the eager step (after the same command exited 0 without ncu)
#   4. `ncu --set full` of the trunk GEMM launches and of one eager step's SIMT / row kernels
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest_gpu_final.log; tail -8 gpurun_out/pytest_gpu_final.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py --steps 100 --warmup 5 > gpurun_out/bench_final_cfg2.json 2> gpurun_out/bench_final_cfg2.err; cut -c1-200 gpurun_out/bench_final_cfg2.json
timeout 100 python scripts/profile_step.py > gpurun_out/step_profile_final.json 2>/dev/null
if [ "$1" = "ncu" ]; then
  timeout 200 python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-tc-rooflines > gpurun_out/plain.log 2>&1 && timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 600 --csv --log-file gpurun_out/launches_cfg2.csv python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-tc-rooflines > gpurun_out/ncu_launch.log 2>&1
  timeout 120 python scripts/ncu_x2.py && timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_x2_gemm --launch-skip 5 --launch-count 5 -f -o gpurun_out/x2 python scripts/ncu_x2.py > gpurun_out/ncu_x2.log 2>&1
  timeout 100 python scripts/ncu_step.py && timeout 400 ncu --set full --clock-control none --import-source on -k 'regex:k_gyro|k_mobius|k_reduce_slabs|k_hradius|k_absmax|k_split2h|k_bce|k_rn_head' --launch-skip 80 --launch-count 28 -f -o gpurun_out/step_kernels python scripts/ncu_step.py > gpurun_out/ncu_step.log 2>&1
fi
