#!/bin/bash
# A/B on ONE box: the build in variants/base (git archive of an earlier commit) against the working tree, interleaved
# usage: tools/ab.sh [rounds] [bench args...]
n=${1:-3}; shift
one() { python bench.py --no-cpu-baseline --steps 192 --warmup 8 "$@" 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.4f' % d['ms_per_step'])"; }
for i in $(seq $n); do
  a=$(GGB_PKG_DIR=$PWD/variants/base/gemma.ggml_b200 one "$@"); b=$(one "$@")
  echo "base $a ms   new $b ms"
done | tee gpurun_out/ab.log
