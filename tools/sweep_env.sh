mkdir -p gpurun_out
# producer knobs of the decode kernel: L2 look-ahead chunks, ring copies in flight
for cfg in "0 2" "0 3" "4 2" "8 3"; do set -- $cfg
  GGB_PREFETCH_CHUNKS=$1 GGB_MAX_OUTSTANDING=$2 timeout 300 python bench.py --no-sub-records --no-cpu-baseline ${BENCH_ARGS:-} > gpurun_out/sw.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/sw.json')); print('prefetch $1 outstanding $2 :', round(1000*d['ms_per_step'],1), 'us')"
done
