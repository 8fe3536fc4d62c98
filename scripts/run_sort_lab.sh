#!/bin/bash
# Runs every sort_lab variant under scripts/build/ (each bounded by its own timeout) and collects the output.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-100000000}
: > gpurun_out/sort_lab.log
for b in scripts/build/sort_lab_*; do
  echo "=== $b" >> gpurun_out/sort_lab.log
  timeout 120 $b $N 3 >> gpurun_out/sort_lab.log 2>&1
  echo "rc=$?" >> gpurun_out/sort_lab.log
done
cat gpurun_out/sort_lab.log
