mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 10 --warmup 3 2>&1 | grep '^{' | tail -1 > gpurun_out/bench_n8.json
python -c "
import json; d=json.load(open('gpurun_out/bench_n8.json')); print(d['value'], d['c3_hunyuan_attn'])"
