#!/bin/bash
# quick GPU check: parity tests + short bench summary.  usage: tools/quick_gpu.sh <tag>
TAG=${1:-q}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -${2:-15}
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -3 gpurun_out/bench_$TAG.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$TAG.json"))
print(round(d["value"]), round(d["ms_per_step"],2), d["clocks"])
print({k:round(v["ms"],2) for k,v in d["legs"].items()})
print({k:round(v["ms_per_step"],2) for k,v in d["roofline"]["per_entry_point"].items()})
PY
