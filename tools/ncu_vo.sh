#!/bin/bash
# ncu captures of the kernels of the VO step (run under gpurun, one GPU): full sets of the top kernels + the launch list.
# usage: tools/ncu_vo.sh <tag>
set -u
tag=${1:-r2}
mkdir -p gpurun_out
cmd="python bench.py --steps 3 --warmup 3 --no-secondary"
for k in local_ba2_kernel pose_only_kernel sparse_align_kernel project_align_kernel track_project_kernel; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 2 -f -o gpurun_out/${tag}_$k $cmd > gpurun_out/${tag}_ncu_$k.log 2>&1
done
