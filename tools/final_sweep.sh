o=gpurun_out
timeout 200 python bench.py --model 7b --no-cpu-baseline --no-sub-records --steps 64 2>/dev/null | tail -1 > $o/m_bench_7b_q4_0.json
for c in 1024 2048 4096 8192; do timeout 200 python bench.py --model 7b --wtype q8_0 --ctx $c --no-cpu-baseline --no-sub-records --steps 48 2>/dev/null | tail -1 > $o/m_bench_7b_q8_0_ctx$c.json; done
timeout 120 python tools_trace.py 2b 2048 > $o/m_trace_2b_q4_0.txt 2>&1
timeout 120 python tools_trace.py 7b 2048 > $o/m_trace_7b_q4_0.txt 2>&1
for f in $o/m_bench_7b_*.json; do python -c "import json,sys; d=json.load(open('$f')); print('$f', round(d.get('value'),1), round(d.get('ms_per_step'),4), round((d.get('roofline') or {}).get('frac'),3))"; done
grep -A7 "layer 5 detail" $o/m_trace_2b_q4_0.txt
