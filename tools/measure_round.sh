#!/bin/bash
# End-of-round measurement on ONE B200 (all outputs under gpurun_out/m_*): tests, bench lines, launch list, ncu capture, traces
set -u
o=gpurun_out; mkdir -p $o
timeout 1500 python -m pytest tests -m gpu -x -q > $o/m_pytest_gpu.log 2>&1; tail -2 $o/m_pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > $o/m_smoke.log 2>&1; tail -1 $o/m_smoke.log
timeout 900 python bench.py 2>$o/m_bench.err | tail -1 > $o/m_bench_2b_q4_0.json
timeout 600 python bench.py --impl reference --steps 8 --warmup 3 2>/dev/null | tail -1 > $o/m_bench_reference.json
timeout 600 python bench.py --wtype q8_0 --no-cpu-baseline --no-sub-records 2>/dev/null | tail -1 > $o/m_bench_2b_q8_0.json
timeout 600 python bench.py --model 7b --no-cpu-baseline --no-sub-records --steps 64 2>/dev/null | tail -1 > $o/m_bench_7b_q4_0.json
for c in 1024 2048 4096 8192; do timeout 600 python bench.py --model 7b --wtype q8_0 --ctx $c --no-cpu-baseline --no-sub-records --steps 48 2>/dev/null | tail -1 > $o/m_bench_7b_q8_0_ctx$c.json; done
timeout 600 python bench.py --mode prefill --wtype q8_0 --steps 5 --no-cpu-baseline --no-sub-records 2>/dev/null | tail -1 > $o/m_bench_prefill_2b_q8_0.json
timeout 600 python bench.py --mode ops --no-cpu-baseline --no-sub-records 2>/dev/null | tail -1 > $o/m_bench_2b_q4_0_ops.json
timeout 120 python tools_trace.py 2b 2048 > $o/m_trace_2b_q4_0.txt 2>&1
timeout 120 python tools_trace.py 7b 2048 > $o/m_trace_7b_q4_0.txt 2>&1
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sub-records > /dev/null 2>&1 && \
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/m_decode_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sub-records > $o/m_ncu_launch.log 2>&1
timeout 300 python tools/run_fused.py 6 > $o/m_plain.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:decoder_kernel -s 3 -c 1 -o $o/m_decode_full -f python tools/run_fused.py 6 > $o/m_ncu_full.log 2>&1
for f in $o/m_bench_*.json; do python -c "import json,sys; d=json.load(open('$f')); print('$f', d.get('value'), d.get('ms_per_step'), (d.get('roofline') or {}).get('frac'), (d.get('e2e') or {}).get('value'))"; done
