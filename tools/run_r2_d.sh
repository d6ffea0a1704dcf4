mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py -x -q > gpurun_out/r2_model_test.log 2>&1; echo "model rc=$?"; tail -3 gpurun_out/r2_model_test.log
timeout 300 python bench.py --no-sub-records > gpurun_out/r2_bench_d.json 2> gpurun_out/r2_bench_d.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_d.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'])"
for ph in 35; do GGB_PROF_PHASE=$ph timeout 120 python tools_trace.py 2b 2048 > gpurun_out/r2_trace_ph$ph.log 2>&1; echo "trace $ph rc=$?"; grep -B12 -A12 "inside phase" gpurun_out/r2_trace_ph$ph.log | head -40; done
