#!/bin/bash
# Round-1 evidence run: GPU tests, bench line, ncu launch list of the bench command, full captures of the three hot kernels.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee gpurun_out/p_pytest.txt
timeout 900 python bench.py --steps 20 --warmup 5 2>gpurun_out/p_bench.err | tail -1 > gpurun_out/p_bench_n1.json
cat gpurun_out/p_bench_n1.json | cut -c1-600
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/p_bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 -k regex:"attn_kernel|mlp_kernel" --csv --log-file gpurun_out/p_launches.csv python bench.py --steps 4 --warmup 3 --no-extras > gpurun_out/p_launch_bench.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"attn_kernel|mlp_kernel" -s 9 -c 3 -o gpurun_out/p_c2 -f python bench.py --steps 4 --warmup 3 --no-extras > gpurun_out/p_ncu_c2.log 2>&1
tail -2 gpurun_out/p_ncu_c2.log
timeout 600 python tools/sample_schedule.py --out gpurun_out/p_sample_schedule.json 2>&1 | tail -2
timeout 900 ncu --set full --import-source on --clock-control none -k regex:attn_kernel -s 1 -c 1 -o gpurun_out/p_c3 -f python tools/quick_attn.py 119056 8320 1 24 1 > gpurun_out/p_ncu_c3.log 2>&1
tail -2 gpurun_out/p_ncu_c3.log
timeout 600 python tools/sweep.py --out gpurun_out/p_sweep.jsonl > /dev/null 2>&1; wc -l gpurun_out/p_sweep.jsonl
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
ls -la gpurun_out
