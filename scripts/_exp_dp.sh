for b in 32 64 128; do
echo "blocks=$b"; HVAE_AR_BLOCKS=$b timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 scripts/dp_probe.py 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print({k:(round(v,1) if isinstance(v,float) else v) for k,v in d.items() if 'us' in k or 'check' in k})"
done
