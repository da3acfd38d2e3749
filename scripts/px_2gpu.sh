#!/bin/bash
# peer-push exchange bring-up on N GPUs: sharded parity, then bench lines for a few settings (run under gpurun --gpus N)
N=${1:-2}
T="timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
mkdir -p gpurun_out
$T scripts/check_sharded.py > gpurun_out/px${N}_check.log 2>&1; echo "rc=$?" >> gpurun_out/px${N}_check.log
tail -3 gpurun_out/px${N}_check.log
IFS=';' read -ra CFGS <<< "${CFGS:-16 0.1 peer 1;16 0.01 peer 1;32 0.1 peer 1;8 0.01 peer 1;16 0.01 peer 0}"
for cfg in "${CFGS[@]}"; do
  set -- $cfg
  KTN_PUSH_BLOCKS=$1 KTN_EXCHANGE=$3 KTN_PUSH_RESERVE=$4 $T bench.py --gpus $N --steps ${STEPS:-100} --warmup 10 --violated $2 > gpurun_out/px${N}_b$1_v$2_$3_r$4.log 2>&1
  echo "cfg $cfg rc=$?"; grep '^{' gpurun_out/px${N}_b$1_v$2_$3_r$4.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['config']['exchange'][:12], d['roofline']['ms_per_launch'], d['roofline']['round']['k2_compact_ms'])"
done
