#!/bin/bash
# usage: scripts/scale_run.sh N  -- the driver's launch line for N GPUs, then exchange sweeps (value leg only)
N=$1; PORT=29517
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $N "$@" 2>/dev/null | grep '^{' | tail -1; PORT=$((PORT+1)); }
echo "== full bench N=$N"; run --steps 100 --warmup 5 | tee gpurun_out/bench_n$N.json | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['round']['k2_k3_ms'], d['e2e']['value'], d['e2e']['ms_per_step'], d['sharded'])"
for b in 8 24 32; do echo "== KTN_PUSH_BLOCKS=$b"; KTN_PUSH_BLOCKS=$b run --steps 100 --warmup 5 --skip-e2e; done
echo "== KTN_PUSH_RESERVE=0"; KTN_PUSH_RESERVE=0 run --steps 100 --warmup 5 --skip-e2e
echo "== nccl"; KTN_EXCHANGE=nccl run --steps 100 --warmup 5 --skip-e2e
echo "== v=0.01"; run --steps 100 --warmup 5 --skip-e2e --violated 0.01
