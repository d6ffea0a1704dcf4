mkdir -p gpurun_out
for ph in ${PHASES:-34}; do GGB_PROF_PHASE=$ph timeout 120 python tools_trace.py ${MODEL:-2b} 2048 > gpurun_out/q_trace_ph$ph.log 2>&1; echo "trace $ph rc=$?"; grep -A40 "^phase " gpurun_out/q_trace_ph$ph.log | head -60; done
