for c in 0 6 12; do echo "== prefetch $c"; GGB_PREFETCH_CHUNKS=$c timeout 120 python tools_trace.py 2b 2048 2>&1 | grep -B1 -A8 "^phase "; done
