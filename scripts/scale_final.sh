#!/bin/bash
# usage: scripts/scale_final.sh N  -- the driver's launch line for N GPUs (default steps), plus the v = 0.01 value leg
N=$1; PORT=29617
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $N "$@" 2>gpurun_out/bench_n$N.err | grep '^{' | tail -1; PORT=$((PORT+1)); }
run | tee gpurun_out/bench_n$N.json | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('N', d['n_gpus'], 'value', d['value'], 'ms', d['ms_per_step'], 'K1', d['roofline']['ms_per_launch'], 'K2K3', d['roofline']['round']['k2_k3_ms'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'parity', d.get('sharded_parity')); print(d['sharded'])"
run --skip-e2e --violated 0.01 | tee gpurun_out/bench_n${N}_v001.json | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('v=0.01 value', d['value'], 'ms', d['ms_per_step'], d['sharded'].get('exchange_ms'))"
tail -2 gpurun_out/bench_n$N.err
