#!/bin/bash
# End-of-round measurement on ONE B200 (all outputs under gpurun_out/m_*): tests, bench lines, launch list, ncu capture
set -u
o=gpurun_out
python -m pytest tests -m gpu -x -q > $o/m_pytest_gpu.log 2>&1; tail -2 $o/m_pytest_gpu.log
python bench.py 2>$o/m_bench.err | tail -1 > $o/m_bench_decode.json
python bench.py --impl reference --steps 8 --warmup 1 2>/dev/null | tail -1 > $o/m_bench_reference.json
python bench.py --wtype q8_0 --no-cpu-baseline 2>/dev/null | tail -1 > $o/m_bench_decode_2b_q8.json
python bench.py --model 7b --no-cpu-baseline --steps 64 2>/dev/null | tail -1 > $o/m_bench_decode_7b_q4.json
for c in 1024 2048 4096 8192; do python bench.py --model 7b --wtype q8_0 --ctx $c --no-cpu-baseline --steps 48 2>/dev/null | tail -1 > $o/m_bench_decode_7b_q8_ctx$c.json; done
python bench.py --mode prefill --wtype q8_0 --steps 5 2>/dev/null | tail -1 > $o/m_bench_prefill_q8.json
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/m_decode_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $o/m_ncu_launch.log 2>&1
python tools/run_fused.py 6 > $o/m_plain.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:decoder_kernel -s 3 -c 1 -o $o/m_decode_full -f python tools/run_fused.py 6 > $o/m_ncu_full.log 2>&1
for f in $o/m_bench_*.json; do python -c "import json,sys; d=json.load(open('$f')); print('$f', d.get('value'), d.get('ms_per_step'), (d.get('roofline') or {}).get('frac'))"; done
