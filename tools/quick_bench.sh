#!/bin/bash
# usage: tools/quick_bench.sh TAG [ENV=VAL ...] -- short device-resident bench of the C3 step, prints the headline numbers
tag=$1; shift
env "$@" python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-collisions --no-c4 --no-mass-matrix ${SORT_EVERY:+--sort-every $SORT_EVERY} > gpurun_out/qb_$tag.json 2>gpurun_out/qb_$tag.err || tail -3 gpurun_out/qb_$tag.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/qb_$tag.json').read().strip().splitlines()[-1])
    print('$tag', '%.3e'%d['value'], 'ms/step', round(d['ms_per_step'],2), 'kern', round(d['roofline']['kernel_ms_per_launch'],4), 'frac', round(d['roofline']['frac'],3), 'defer', d['kernel_ms_per_step'].get('advance_deferred'), 'k', d['config']['mean_picard_passes'], 'unconv', d['config']['unconverged_particles'])
except Exception as e: print('$tag ERR', e)
PY
