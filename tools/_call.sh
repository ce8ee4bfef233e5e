mkdir -p gpurun_out
for n in 4 8; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 2>&1 | grep '^{' | tail -1 > gpurun_out/bench_n$n.json
cut -c1-300 gpurun_out/bench_n$n.json
done
