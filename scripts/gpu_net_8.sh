#!/bin/bash
# 8-GPU (or N-GPU) network leg: bulk push (default), per-thread stores (-DJDN_BULK_PUSH=0), phase timers
N=${1:-8}; NODES=${2:-1000000}
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 scripts/gpu_net_scale.py $NODES 2>&1; }
echo "== bulk push"; run 29601 | grep "^{" | tail -1
echo "== per-thread stores"; JITCDDE_B200_NET_FLAGS=-DJDN_BULK_PUSH=0 run 29602 | grep "^{" | tail -1
echo "== bulk push again"; run 29603 | grep "^{" | tail -1
echo "== phases (bulk push)"; JITCDDE_B200_NET_FLAGS=-DJDN_PROFILE run 29604 > gpurun_out/r2_net_phases_${N}_bulk.log 2>&1; grep -o "rank 0 attempts 5[0-9].*" gpurun_out/r2_net_phases_${N}_bulk.log | head -2 | cut -c1-600
