mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py -x -q > gpurun_out/q_model_test.log 2>&1; echo "model rc=$?"; tail -2 gpurun_out/q_model_test.log
timeout 300 python bench.py --no-sub-records --no-cpu-baseline > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err; python -c "
import json; d=json.load(open('gpurun_out/q_bench.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'])"
