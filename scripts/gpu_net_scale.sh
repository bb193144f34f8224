#!/bin/bash
# usage: scripts/gpu_net_scale.sh MAXGPUS [nodes] — network leg at 1, 2, 4, 8 GPUs
MAX=${1:-8}; NODES=${2:-1000000}
for g in 1 2 4 8; do
	if [ "$g" -le "$MAX" ]; then
		if [ "$g" -eq 1 ]; then python scripts/gpu_net_scale.py $NODES 2>&1 | tail -1
		else python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 2951$g scripts/gpu_net_scale.py $NODES 2>&1 | grep "^{" | tail -1; fi
	fi
done
