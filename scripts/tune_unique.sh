# range (?a) hash pass: geometry / prefetch-depth knobs, each built on the box and timed with probe_unique.py
# usage (under gpurun): bash scripts/tune_unique.sh "CT U" ...
out=gpurun_out/${TUNE_OUT:-tune_unique.log}
for cfg in "$@"; do set -- $cfg
  KGB200_NVCC_DEFS="-DKGB_UNIQ_CT=$1 -DKGB_UNIQ_CU=$2" python -m klongpy_b200.csrc.build --force >/dev/null 2>&1
  echo "CT=$1 U=$2" >> $out
  python scripts/probe_unique.py 2>&1 | grep -v "distinct (fallback)\|permutation\|500k" >> $out
done
cat $out
