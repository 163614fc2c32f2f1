#!/bin/bash
# usage: tools/bench_env_sweep.sh "VAR=a VAR2=b" "VAR=c" ...  -- one short bench per env setting
for cfg in "$@"; do
  env $cfg python bench.py --steps 5 --warmup 2 --no-cpu-baseline --no-e2e 2> gpurun_out/b.err | grep -v '^#' > gpurun_out/b.json
  python - "$cfg" <<'PY'
import json, sys
try:
    d = json.load(open("gpurun_out/b.json"))
    print("%-50s value %.4g  step %.2f ms  frac %.3f  kernel %.3f ms  k=%s unconv=%s" % (sys.argv[1], d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["kernel_ms_per_launch"], d["config"]["mean_picard_passes"], d["config"]["unconverged_particles"]))
except Exception as e:
    print(sys.argv[1], "FAILED", e); print(open("gpurun_out/b.err").read()[-1500:])
PY
done
