set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest_gpu_final.log; tail -8 gpurun_out/pytest_gpu_final.log
cp gpurun_out/parity_audit_*.json gpurun_out/parity_audit_final.json 2>/dev/null
timeout 100 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 300 python bench.py --steps 100 --warmup 5 > gpurun_out/bench_final_cfg2.json 2> gpurun_out/bench_final_cfg2.err; cut -c1-200 gpurun_out/bench_final_cfg2.json
timeout 100 python scripts/profile_step.py > gpurun_out/step_profile_final.json 2>/dev/null
