#!/bin/bash
# A/B/n on ONE box: builds under variants/<name>/ (copies of the package built from other sources) and the working
# tree ("."), interleaved.  usage: tools/ab.sh <rounds> <variant> [<variant> ...] [-- bench args]
n=${1:-3}; shift
vars=(); while [ $# -gt 0 ] && [ "$1" != "--" ]; do vars+=("$1"); shift; done; [ "$1" = "--" ] && shift
[ ${#vars[@]} -eq 0 ] && vars=(base .)
one() { python bench.py --no-cpu-baseline --steps 192 --warmup 8 "$@" 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.4f' % d['ms_per_step'])"; }
for i in $(seq $n); do
  line=""
  for v in "${vars[@]}"; do
    if [ "$v" = "." ]; then r=$(one "$@"); else r=$(GGB_PKG_DIR=$PWD/variants/$v/gemma.ggml_b200 one "$@"); fi
    line="$line $v=$r"
  done
  echo "$line"
done | tee gpurun_out/ab.log
