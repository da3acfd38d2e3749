#!/bin/bash
# 8 GPUs: the driver's launch line, the v = 0.01 value leg, and two push-grid variants around the default (32 blocks)
bash scripts/scale_final.sh 8
PORT=29717
for b in 24 40; do
  KTN_PUSH_BLOCKS=$b python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus 8 --skip-e2e 2>/dev/null | grep '^{' | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('push blocks $b: value', d['value'], 'ms', d['ms_per_step'], d['sharded'].get('exchange_ms'))"
  PORT=$((PORT+1))
done
