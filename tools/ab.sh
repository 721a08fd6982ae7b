#!/bin/bash
# A/B of library variants on the GPU box: tools/ab.sh OUT_PREFIX NAME[=path] ...   ("default" = the in-tree library)
# Writes gpurun_out/OUT_PREFIX_NAME.json (bench line) and gpurun_out/OUT_PREFIX_NAME.micro (config-2 microbench).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
prefix=$1; shift
for v in "$@"; do
  if [ "$v" = default ]; then unset DIFFRP_B200_LIB; else export DIFFRP_B200_LIB=$PWD/build/variants/$v.so; fi
  DIFFRP_B200_NO_BUILD=1 timeout 600 python bench.py --steps 64 --warmup 3 --no-torch-baseline --no-cpu-baseline --no-strong --e2e-steps 1 \
      > gpurun_out/${prefix}_$v.json 2> gpurun_out/${prefix}_$v.err
  DIFFRP_B200_NO_BUILD=1 timeout 300 python tools/microbench_trace.py > gpurun_out/${prefix}_$v.micro 2>&1
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${prefix}_$v.json").read().strip().splitlines()[-1])
    print("$v", "ms/step", d.get("ms_per_step"), "value", d.get("value"), "roofline", d.get("roofline"))
except Exception as e:
    print("$v", "failed", e)
PY
  tail -2 gpurun_out/${prefix}_$v.micro
done
