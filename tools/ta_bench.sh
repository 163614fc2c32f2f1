python - <<'PY'
import json,subprocess,sys
out=subprocess.run([sys.executable,"bench.py","--steps","3","--warmup","3","--no-cpu-baseline","--no-e2e","--no-c4","--no-mass-matrix"],capture_output=True,text=True)
try:
    d=json.loads(out.stdout.strip().splitlines()[-1]); c=d["collisions"]
    print("TA pairs/s %.3e (kernels only %.3e) ms/step %.3f kernels %.3f frac %.3f"%(c["value"],c["value_kernels_only"],c["ms_per_step"],c["ms_per_step_kernels_only"],c["roofline"]["frac"])); print(c["prep_kernel_ms_per_step"])
except Exception as e:
    print("ERR",e,out.stderr[-2000:])
PY
