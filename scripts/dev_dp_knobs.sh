#!/bin/bash
# Data-parallel step time under NCCL knobs (scripts/dev_dp_knobs.sh N): which setting costs the backward pass least.
N=${1:-2}
run() {
  echo "=== $*"
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) \
      bench.py --gpus $N --steps 8 --warmup 3 --no-kernel-leg 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms/step %.2f  img/s %.0f  clocks %s' % (d['ms_per_step'], d['value'], d['clocks']['sm_mhz']))"
}
run VAW_X=0
run NCCL_MAX_CTAS=4
run NCCL_MAX_CTAS=8
run NCCL_MAX_CTAS=16
run NCCL_MIN_CTAS=32
run NCCL_ALGO=Ring
run NCCL_NVLS_ENABLE=0
run VAW_DP_NOSYNC=1
