#!/bin/bash
# One GPU-box visit: parity tests, bench line, ncu launch list, ncu --set full of the layer kernels.
#   tools/gpu_round.sh <tag> [tests|notests] [kernel-regex]
# Everything lands in gpurun_out/ with the tag in the name; copy what should be judged into profiles/.
set -u
TAG=${1:-x}
TESTS=${2:-tests}
KREGEX=${3:-'k_gather|k_gcn_fwd_tc|k_sage_fwd_gemm|k_gcn_bwd_gemm|k_sage_bwd'}
OUT=gpurun_out
mkdir -p $OUT
if [ "$TESTS" = "tests" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $OUT/gpu_tests_$TAG.log 2>&1
  echo "pytest rc=$?" | tee -a $OUT/gpu_tests_$TAG.log
  grep -E "^(FAILED|ERROR|E  )" $OUT/gpu_tests_$TAG.log | head -20
  tail -3 $OUT/gpu_tests_$TAG.log
fi
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
echo "bench rc=$?"
cut -c1-400 $OUT/bench_$TAG.json
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 300 $CMD > $OUT/plain_$TAG.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 700 --csv \
    --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_launch_$TAG.log 2>&1
echo "launch list rc=$?"
# full captures: matched launches of one step are, in order,
#   gcn train: fwd_tc x3, (gather, gemm) x3 | sage train: (gather, gemm) x3, bwd kernels | gcn infer | sage infer
# 29 per step, 3 warm-up steps + 3 ... = skip 87; keep each report small (<= 64 MiB comes back in total)
i=0
for SPEC in ${SPECS:-"89:5" "98:2"}; do
  S=${SPEC%%:*}; C=${SPEC##*:}; i=$((i+1))
  timeout 300 $CMD > $OUT/plain2_$TAG.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$KREGEX" -s $S -c $C \
      -f -o $OUT/prof_${TAG}_$i $CMD > $OUT/ncu_full_${TAG}_$i.log 2>&1
  echo "full capture $i ($SPEC) rc=$?"
  # the summary always travels; the report itself only while the whole directory stays under the 64 MiB that come back
  ncu -i $OUT/prof_${TAG}_$i.ncu-rep --page raw --csv > $OUT/prof_${TAG}_$i.raw.csv 2>/dev/null
  if [ $(du -sm $OUT | cut -f1) -gt 56 ]; then rm -f $OUT/prof_${TAG}_$i.ncu-rep; echo "dropped prof_${TAG}_$i.ncu-rep (size)"; fi
done
du -sh $OUT; ls -la $OUT | tail -14
