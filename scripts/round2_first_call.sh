#!/bin/bash
# First GPU call of the next round: measures the experiments prepared (but not run) at the end of round 1.
# Here (no GPU):   bash scripts/round2_first_call.sh build
# On the box:      gpurun --timeout 900 -- 'bash scripts/round2_first_call.sh run'      (one GPU)
#                  gpurun --gpus 2 --timeout 600 -- 'bash scripts/round2_first_call.sh run2'   (exchange variants)
set -e
cd "$(dirname "$0")/.."
case "$1" in
build)
  bash scripts/build_variants.sh base "" \
       fwd3g4 "-DKTN_FWD_BLOCKS=3 -DKTN_FWD_GROUP=4" fwd1 "-DKTN_FWD_BLOCKS=1" timing "-DKTN_OPT_TIMING" pushtma "-DKTN_OPT_PUSH_TMA"
  ;;
run)
  mkdir -p gpurun_out
  # one process per variant (the binding loads the library RTLD_GLOBAL); equal digests = equal results
  for v in base fwd3g4 fwd1 timing; do KTN_DEBUG=1 timeout 200 python scripts/ab_time.py build/variants/libktn_$v.so 2>&1 | tee -a gpurun_out/ab_round2.log; done
  ;;
run2)
  mkdir -p gpurun_out
  N=${2:-2}
  T="timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
  $T scripts/check_sharded.py 2>&1 | tail -3 | tee -a gpurun_out/run2.log          # includes the sharded top-k check
  for v in base pushtma; do for b in 4 8 16; do
    KTN_LIB=build/variants/libktn_$v.so KTN_PUSH_BLOCKS=$b $T bench.py --gpus $N --steps 200 --warmup 10 2>/dev/null | grep '^{' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v', $b, d['value'], d['ms_per_step'], d['roofline']['ms_per_launch'])" | tee -a gpurun_out/run2.log
  done; done
  ;;
*) echo "usage: $0 build|run|run2 [ngpus]"; exit 2;;
esac
