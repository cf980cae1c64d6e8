HVAE_DP_P2P=1 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29547 bench.py --gpus 8 --steps 50 --warmup 5 --no-cpu-baseline 2> gpurun_out/bench_n8.err | tail -1 > gpurun_out/bench_n8.json
tail -c 400 gpurun_out/bench_n8.err | tail -3
python -c "
import json; d=json.load(open('gpurun_out/bench_n8.json')); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])"
