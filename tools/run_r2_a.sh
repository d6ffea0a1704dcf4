set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kquants_gpu.py -x -q > gpurun_out/r2_kq_test.log 2>&1; echo "kq rc=$?"
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py > gpurun_out/r2_bench_q4.json 2> gpurun_out/r2_bench_q4.err; echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-sub-records > gpurun_out/r2_ncu_launch.log 2>&1; echo "ncu1 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decoder_kernel -c 1 -o gpurun_out/r2_decoder python bench.py --steps 2 --warmup 3 --no-sub-records > gpurun_out/r2_ncu_full.log 2>&1; echo "ncu2 rc=$?"
tail -3 gpurun_out/r2_kq_test.log; tail -2 gpurun_out/r2_smoke.log; cat gpurun_out/r2_bench_q4.json | head -c 3000
