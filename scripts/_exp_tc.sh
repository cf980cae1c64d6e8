timeout 300 python -m pytest tests/test_gpu_tc.py -m gpu -q -x 2>&1 | tail -5
for dbg in 0 1; do
  echo "dbg=$dbg"; HVAE_TC_DBG=$dbg timeout 200 python scripts/microbench_kernels.py --logB 18 --only tc 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print({k:(round(v['ms'],3)) for k,v in d.items() if 'ms' in v})"
done
