#!/bin/bash
# Runs every sort_lab variant under scripts/build/ (each bounded by its own timeout) and collects the output.
#   run_sort_lab.sh [n] [stagger values...]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-100000000}
shift
STAG=${@:-"30"}
: > gpurun_out/sort_lab.log
for b in scripts/build/sort_lab_*; do
  for s in $STAG; do
    echo "=== $b stagger=$s" >> gpurun_out/sort_lab.log
    KGB200_SORT_STAGGER_NS=$s timeout 120 $b $N 3 >> gpurun_out/sort_lab.log 2>&1
    echo "rc=$?" >> gpurun_out/sort_lab.log
  done
done
grep -v "^rand40\|^dup3" gpurun_out/sort_lab.log
