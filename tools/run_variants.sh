# usage: bash tools/run_variants.sh name1 name2 ...  (benches variants/<name>/libggb_cuda.so in place of the in-tree library)
mkdir -p gpurun_out
cp gemma.ggml_b200/libggb_cuda.so /tmp/libggb_cuda.base.so
for v in base "$@"; do
  if [ "$v" = base ]; then cp /tmp/libggb_cuda.base.so gemma.ggml_b200/libggb_cuda.so; else cp variants/$v/libggb_cuda.so gemma.ggml_b200/libggb_cuda.so; fi
  timeout 300 python bench.py --no-sub-records > gpurun_out/var_$v.json 2> gpurun_out/var_$v.err; echo "$v bench rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/var_$v.json')); print('$v', 'us/token', round(1000*d['ms_per_step'],1))"
  timeout 120 python tools_trace.py 2b 2048 > gpurun_out/var_trace_$v.log 2>&1; grep -A7 "layer 5 detail" gpurun_out/var_trace_$v.log
done
cp /tmp/libggb_cuda.base.so gemma.ggml_b200/libggb_cuda.so
