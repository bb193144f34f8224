#!/bin/bash
# N-GPU network leg twice + phase timers (short)
N=${1:-8}; NODES=${2:-1000000}
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 scripts/gpu_net_scale.py $NODES 2>&1; }
run 29611 | grep "^{" | tail -1
JITCDDE_B200_NET_FLAGS=-DJDN_PROFILE run 29614 > gpurun_out/r2_net_phases_${N}_relaxed.log 2>&1
