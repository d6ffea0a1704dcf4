#!/bin/bash
# tensor-parallel measurement on N GPUs of one box: parity test for world = N, then the fused decode bench (2B, 7B)
N=$1; o=gpurun_out
timeout 300 python -m pytest tests/test_tp_gpu.py -x -q -k "$N" > $o/m_tp${N}_pytest.log 2>&1; tail -2 $o/m_tp${N}_pytest.log
for m in 2b 7b; do
  GGB_TP_FUSED=1 timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$N bench.py --gpus $N --steps 128 --warmup 8 --model $m --no-cpu-baseline 2>$o/m_tp${N}_$m.err | tail -1 > $o/m_tp${N}_${m}_fused.json
  python -c "import json; d=json.load(open('$o/m_tp${N}_${m}_fused.json')); print('tp$N $m', d['value'], d['ms_per_step'], d['roofline']['frac'])"
done
