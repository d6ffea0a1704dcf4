mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kquants_gpu.py -x -q > gpurun_out/r2_kq_test.log 2>&1; echo "kq rc=$?"; tail -3 gpurun_out/r2_kq_test.log
for ph in 35 34 30 33; do GGB_PROF_PHASE=$ph timeout 120 python tools_trace.py 2b 2048 > gpurun_out/r2_trace_ph$ph.log 2>&1; echo "trace $ph rc=$?"; done
cat gpurun_out/r2_trace_ph35.log
