mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_model_gpu.py -x -q > gpurun_out/q_model_test.log 2>&1; echo "model rc=$?"; tail -2 gpurun_out/q_model_test.log
for args in "" "--model 7b --steps 64" "--wtype q8_0" "--model 7b --wtype q8_0 --ctx 8192 --steps 32"; do
timeout 200 python bench.py --no-sub-records --no-cpu-baseline $args > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err; python -c "
import json; d=json.load(open('gpurun_out/q_bench.json')); print('$args', round(d['value'],1), round(1000*d['ms_per_step'],1), round(d['roofline']['frac'],3))"
done
