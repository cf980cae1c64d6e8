set -x
timeout 400 python bench.py --steps 100 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 400 gpurun_out/bench_n1.err
timeout 300 python scripts/microbench_kernels.py --logB 18 > gpurun_out/microbench.json 2> gpurun_out/microbench.err
timeout 200 python scripts/profile_step.py > gpurun_out/step_profile.json 2> gpurun_out/step_profile.err
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/plain.log 2>&1 && timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 600 --csv --log-file gpurun_out/launches_r1d.csv python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu.log 2>&1
tail -1 gpurun_out/ncu.log | cut -c1-120
