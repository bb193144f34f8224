#!/bin/bash
# usage: scripts/gpu_ncu_mg.sh <tag> "<variant spec>" [launch-skip]  — ncu --set full of one jdde_k_integrate launch of the bench workload
tag=$1; spec=$2; skip=${3:-16}
WARM=12 STEPS=8 python scripts/gpu_variants.py "$spec" > gpurun_out/${tag}_plain.log 2>&1 || exit 1
WARM=12 STEPS=8 ncu --set full --clock-control none --import-source on -k regex:jdde_k_integrate -s $skip -c 1 -f -o gpurun_out/$tag python scripts/gpu_variants.py "$spec" > gpurun_out/${tag}_ncu.log 2>&1
tail -2 gpurun_out/${tag}_plain.log
