#!/bin/bash
# Runs on the GPU box under gpurun: tests, bench, ncu launch list, per-kernel ncu full captures (kept small).
# usage: scripts/gpu_call.sh <tag> [steps...]   steps: test bench ref launches ncu:<only>:<regex> ...
tag=$1; shift
out=gpurun_out; mkdir -p $out
for step in "$@"; do
  case $step in
    test)
      timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest exit=$?" | tee -a $out/${tag}_pytest_gpu.log
      grep -E "^(FAILED|ERROR)|passed|failed" $out/${tag}_pytest_gpu.log | tail -5 ;;
    bench)
      timeout 900 python bench.py --steps ${BENCH_STEPS:-5} --warmup 3 > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench exit=$?"
      tail -c 400 $out/${tag}_bench.err; head -c 1500 $out/${tag}_bench.json ;;
    ref)
      timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err; echo "ref exit=$?" ;;
    launches)
      python bench.py --steps 2 --warmup 3 --no-e2e > $out/${tag}_plain.log 2>&1 && \
      ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_launches.csv \
        python bench.py --steps 2 --warmup 3 --no-e2e > $out/${tag}_ncu_launches.log 2>&1; echo "launches exit=$?" ;;
    ncu:*)
      IFS=: read -r _ only regex <<< "$step"
      python scripts/ncu_targets.py --only $only > $out/${tag}_ncu_plain_$only.log 2>&1 && \
      ncu --set full --clock-control none --import-source on -k regex:"$regex" -c ${NCU_COUNT:-12} -o $out/${tag}_prof_$only \
        python scripts/ncu_targets.py --only $only > $out/${tag}_ncu_$only.log 2>&1; echo "ncu $only exit=$?"
      rep=$out/${tag}_prof_$only.ncu-rep
      if [ -f $rep ]; then
        ncu -i $rep --page raw --csv > $out/${tag}_raw_$only.csv 2>/dev/null
        ncu -i $rep --page details --csv > $out/${tag}_details_$only.csv 2>/dev/null
        ncu -i $rep --page source --csv > $out/${tag}_source_$only.csv 2>/dev/null; gzip -f $out/${tag}_source_$only.csv
        sz=$(du -m $rep | cut -f1); if [ "$sz" -gt 12 ]; then rm -f $rep; fi
      fi ;;
    traffic)
      # DRAM bytes of the dominant kernel (the 1e9-element running sum of bench.py) from one ncu --set full capture
      python bench.py --steps 2 --warmup 3 --no-e2e --no-primitives > $out/${tag}_traffic_plain.log 2>&1 && \
      ncu --set full --clock-control none --import-source on -k regex:kgb_scan -s 3 -c 1 -o $out/${tag}_prof_bench_scan \
        python bench.py --steps 2 --warmup 3 --no-e2e --no-primitives > $out/${tag}_traffic_ncu.log 2>&1; echo "traffic exit=$?"
      rep=$out/${tag}_prof_bench_scan.ncu-rep
      if [ -f $rep ]; then
        ncu -i $rep --page raw --csv > $out/${tag}_raw_bench_scan.csv 2>/dev/null
        ncu -i $rep --page source --csv > $out/${tag}_source_bench_scan.csv 2>/dev/null; gzip -f $out/${tag}_source_bench_scan.csv
        sz=$(du -m $rep | cut -f1); if [ "$sz" -gt 12 ]; then rm -f $rep; fi
      fi ;;
    firstlight)
      timeout 900 python scripts/gpu_first_light.py --perf > $out/${tag}_first_light.log 2>&1; echo "firstlight exit=$?"; grep -E "perf|FAIL" $out/${tag}_first_light.log ;;
    *) echo "unknown step $step" ;;
  esac
done
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
du -sm $out
