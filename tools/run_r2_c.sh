mkdir -p gpurun_out
for ph in 35 32 31; do GGB_PROF_PHASE=$ph timeout 120 python tools_trace.py 2b 2048 > gpurun_out/r2_trace_ph$ph.log 2>&1; echo "trace $ph rc=$?"; grep -A12 "inside phase" gpurun_out/r2_trace_ph$ph.log | head -14; done
