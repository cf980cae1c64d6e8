N=$1
for p2p in 1 0; do
echo "N=$N p2p=$p2p"
HVAE_DP_P2P=$p2p timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$p2p bench.py --gpus $N --steps 50 --warmup 5 --no-cpu-baseline 2> gpurun_out/bench_n${N}_p$p2p.err | tail -1 > gpurun_out/bench_n${N}_p$p2p.json
tail -c 300 gpurun_out/bench_n${N}_p$p2p.err | tail -2
python -c "
import json; d=json.load(open('gpurun_out/bench_n${N}_p$p2p.json')); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'])"
done
