#!/bin/bash
N=8; PORT=29617
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $N "$@" 2>/dev/null | grep '^{' | tail -1; PORT=$((PORT+1)); }
echo "== full bench N=$N"; run --steps 100 --warmup 5 | tee gpurun_out/bench_n$N.json | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['round']['k2_k3_ms'], d['e2e']['value'], d['e2e']['ms_per_step'], d['sharded'])"
for b in 32 48; do echo "== KTN_PUSH_BLOCKS=$b"; KTN_PUSH_BLOCKS=$b run --steps 100 --warmup 5 --skip-e2e; done
echo "== v=0.01"; run --steps 100 --warmup 5 --skip-e2e --violated 0.01
