#!/bin/bash
# Builds one sort_lab binary per pass geometry (CT KPT LBT OCC ATOM LB) in parallel into scripts/build/.
set -e
cd "$(dirname "$0")/.."
mkdir -p scripts/build
VARIANTS=${VARIANTS:-"512,16,128,1,0,4 512,16,128,1,1,4 512,16,256,1,0,4 512,8,64,2,0,4 256,16,64,2,0,4 256,16,64,3,0,4"}
for v in $VARIANTS; do
  IFS=, read ct kpt lbt occ atom lb stats nolb <<< "$v"
  stats=${stats:-0}; nolb=${nolb:-0}
  out=scripts/build/sort_lab_${ct}x${kpt}_l${lbt}_o${occ}_a${atom}_b${lb}_s${stats}_n${nolb}
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -I klongpy_b200/csrc -I include \
    -DKGB_PK_CT=$ct -DKGB_PK_KPT=$kpt -DKGB_PK_LBT=$lbt -DKGB_PK_OCC=$occ -DKGB_PK_ATOM=$atom -DKGB_PK_LB=$lb -DKGB_PK_STATS=$stats -DKGB_PK_NOLB=$nolb \
    scripts/sort_lab.cu -o $out &
done
wait
ls -la scripts/build/
