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
timeout 300 $CMD > $OUT/plain2_$TAG.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$KREGEX" -s 87 -c 29 \
    -f -o $OUT/prof_$TAG $CMD > $OUT/ncu_full_$TAG.log 2>&1
echo "full capture rc=$?"
ls -la $OUT | tail -12
