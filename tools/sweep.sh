#!/bin/bash
# sweep producer tuning knobs of the persistent decode kernel: prints ms/token per setting
out=gpurun_out/sweep.log; : > $out
for pf in ${PF_LIST:-0 4 8 12 16}; do for mo in ${MO_LIST:-2 3}; do
  r=$(GGB_PREFETCH_CHUNKS=$pf GGB_MAX_OUTSTANDING=$mo python bench.py --no-cpu-baseline --steps 128 --warmup 8 ${BENCH_ARGS} 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['roofline']['frac'])")
  echo "prefetch=$pf outstanding=$mo -> $r" | tee -a $out
done; done
