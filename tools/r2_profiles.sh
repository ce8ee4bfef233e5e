#!/bin/bash
# Round-2 evidence run (under gpurun, 1 GPU): GPU tests, bench lines, ncu launch list of the bench command, full ncu
# captures of the hot kernels, index-kernel throughput, the 49-step schedule, the sparsity x sequence sweep, smoke().
# tools/make_profiles.py r02 turns gpurun_out/q_* into the tracked files under profiles/.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -3 | tee gpurun_out/q_pytest.txt
timeout 900 python bench.py --steps 20 --warmup 5 2>gpurun_out/q_bench.err | tail -1 > gpurun_out/q_bench_n1.json
cut -c1-400 gpurun_out/q_bench_n1.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/q_bench_ref.json
M="gpu__time_duration.sum"
timeout 900 ncu --metrics $M --clock-control none -k regex:"attn_kernel|mask_to_indices|mlp_kernel|dense_kernel|select_columns|bitpack|gather_rows" -c 400 --csv --log-file gpurun_out/q_launches.csv python bench.py --steps 4 --warmup 3 --no-extras > gpurun_out/q_launch_bench.log 2>&1
X="--set full --import-source on --clock-control none --metrics l1tex__m_xbar2l1tex_read_bytes.sum,l1tex__m_xbar2l1tex_read_bytes.sum.per_second,l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read.sum.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum.per_second"
timeout 900 ncu $X -k regex:"attn_kernel|mask_to_indices" -s 2 -c 2 -o gpurun_out/q_c3 -f python tools/prof_targets.py c3attn > gpurun_out/q_ncu_c3.log 2>&1; tail -1 gpurun_out/q_ncu_c3.log
timeout 900 ncu $X -k regex:"attn_kernel|mlp_kernel" -s 6 -c 3 -o gpurun_out/q_c2 -f python tools/prof_targets.py c2 > gpurun_out/q_ncu_c2.log 2>&1; tail -1 gpurun_out/q_ncu_c2.log
timeout 900 ncu $X -k regex:dense_kernel -c 1 -o gpurun_out/q_dense_cs -f python tools/prof_targets.py dense_cs > gpurun_out/q_ncu_dense_cs.log 2>&1
timeout 900 ncu $X -k regex:dense_kernel -c 1 -o gpurun_out/q_dense -f python tools/prof_targets.py dense > gpurun_out/q_ncu_dense.log 2>&1
timeout 900 ncu $X -k regex:select_columns -c 1 -o gpurun_out/q_select -f python tools/prof_targets.py select > gpurun_out/q_ncu_select.log 2>&1
timeout 300 python tools/bench_index_kernels.py --out gpurun_out/q_index.json > gpurun_out/q_index.txt 2>&1
timeout 600 python tools/sample_schedule.py --out gpurun_out/q_sample_schedule.json 2>&1 | tail -1 | cut -c1-300
timeout 900 python tools/sweep.py --out gpurun_out/q_sweep.jsonl > /dev/null 2>&1; wc -l gpurun_out/q_sweep.jsonl
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
ls gpurun_out | grep q_
