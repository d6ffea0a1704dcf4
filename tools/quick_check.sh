mkdir -p gpurun_out
timeout 120 python __graft_entry__.py --smoke > gpurun_out/q_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/q_smoke.log
timeout 400 python -m pytest tests/test_model_gpu.py -x -q > gpurun_out/q_model_test.log 2>&1; echo "model rc=$?"; tail -2 gpurun_out/q_model_test.log
timeout 200 python bench.py --no-sub-records --no-cpu-baseline ${BENCH_ARGS:-} > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err; python -c "
import json; d=json.load(open('gpurun_out/q_bench.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'])"
