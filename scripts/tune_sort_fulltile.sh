# grade: default build vs -DKGB_SORT_FULLTILE=1 (built on the box): timing on the three config-3 key distributions
# (tune_sort.py checks every result against torch's stable argsort) and the sort/group parity tests on each build
out=gpurun_out/${TUNE_OUT:-tune_sort_fulltile.log}
for ft in 0 1; do
  KGB200_NVCC_DEFS="-DKGB_SORT_FULLTILE=$ft" python -m klongpy_b200.csrc.build --force >/dev/null 2>&1
  python scripts/tune_sort.py 1e8 fulltile=$ft >> $out 2>&1
  python scripts/tune_sort.py 100000007 fulltile=$ft,ragged >> $out 2>&1
  python -m pytest tests/test_gpu_parity.py tests/test_full_size.py -m gpu -x -q -k "grade or sort or group or argsort or range" 2>&1 | tail -1 >> $out
done
cat $out
