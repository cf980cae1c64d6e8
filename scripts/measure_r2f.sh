set -x
timeout 200 python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-tc-rooflines > gpurun_out/plain.log 2>&1 && timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 600 --csv --log-file gpurun_out/launches_r2f_cfg2.csv python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-tc-rooflines > gpurun_out/ncu_launch.log 2>&1
tail -1 gpurun_out/ncu_launch.log | cut -c1-100
for w in cfg1 cfg1b cfg3 cfg4; do timeout 200 python bench.py --workload $w --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2f_$w.json 2> gpurun_out/bench_r2f_$w.err; cut -c1-160 gpurun_out/bench_r2f_$w.json; done
