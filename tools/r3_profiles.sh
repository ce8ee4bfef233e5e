#!/bin/bash
# Final round-2 evidence run (under gpurun, 1 GPU) after the M=64 / four-threads-per-row change of the attention kernel:
# GPU tests, the bench line, the ncu launch list of the bench command, a full ncu capture of the headline step, smoke().
# Every step has its own timeout.  tools/make_profiles_r2.py turns gpurun_out/q_* into the tracked files under profiles/.
mkdir -p gpurun_out
timeout -k 5 300 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -3 | tee gpurun_out/q_pytest.txt
timeout -k 5 400 python bench.py --steps 20 --warmup 5 2>gpurun_out/q_bench.err | tail -1 > gpurun_out/q_bench_n1.json
cut -c1-300 gpurun_out/q_bench_n1.json
M="gpu__time_duration.sum"
timeout -k 5 200 ncu --metrics $M --clock-control none -k regex:"attn_kernel|mask_to_indices|mlp_kernel|dense_kernel|select_columns|bitpack|gather_rows" -c 400 --csv --log-file gpurun_out/q_launches.csv python bench.py --steps 4 --warmup 3 --no-extras > gpurun_out/q_launch_bench.log 2>&1
X="--set full --import-source on --clock-control none --metrics l1tex__m_xbar2l1tex_read_bytes.sum,l1tex__m_xbar2l1tex_read_bytes.sum.per_second,l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read.sum.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum.per_second"
timeout -k 5 200 ncu $X -k regex:"attn_kernel|mask_to_indices" -s 2 -c 2 -o gpurun_out/q_c3 -f python tools/prof_targets.py c3attn > gpurun_out/q_ncu_c3.log 2>&1; tail -1 gpurun_out/q_ncu_c3.log
timeout -k 5 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
