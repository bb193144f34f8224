#!/bin/bash
# usage: scripts/gpu_net_multi.sh N   — network leg of bench.py on N GPUs (torchrun), plus the two-device peer test when N >= 2
N=${1:-2}
if [ "$N" -ge 2 ]; then
	python -m pytest tests/test_gpu_network.py -q -s -k "fused_peer" 2>&1 | tail -6
fi
for g in 1 2 4 8; do
	if [ "$g" -le "$N" ]; then
		if [ "$g" -eq 1 ]; then
			python bench.py --gpus 1 --no-cpu --no-parity --steps 10 --warmup 3 > gpurun_out/r2_net_${g}gpu.json 2> gpurun_out/r2_net_${g}gpu.err
		else
			python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $g --no-cpu --no-parity --steps 10 --warmup 3 > gpurun_out/r2_net_${g}gpu.json 2> gpurun_out/r2_net_${g}gpu.err
		fi
		tail -2 gpurun_out/r2_net_${g}gpu.err
		python -c "
import json,sys
d=json.loads([l for l in open('gpurun_out/r2_net_${g}gpu.json') if l.startswith('{')][-1])
n=d.get('network',{})
print('gpus', $g, 'value', d['value'], 'net', {k:n.get(k) for k in ('value','ms_per_attempt','accepted','attempts','gpu_launches','checksum','error','exchange')})
"
	fi
done
