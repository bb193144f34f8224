#!/bin/bash
# C3 leg under different Gram-Schmidt team shapes (JITCDDE_B200_TEAM_SMEM bytes per block, JITCDDE_B200_TEAM_THREADS)
N=${1:-100000}
for cfg in "0 512" "113000 256" "113000 512" "75000 256" "56000 256" "113000 384"; do
  set -- $cfg
  echo "== smem $1 threads $2"
  if [ "$1" = 0 ]; then python scripts/gpu_c3_leg.py $N; else JITCDDE_B200_TEAM_SMEM=$1 JITCDDE_B200_TEAM_THREADS=$2 python scripts/gpu_c3_leg.py $N; fi
done
