#!/bin/bash
# usage: tools/ncu_occ.sh TAG ENV=VAL ... : occupancy / utilisation counters of the 2D CC1 tile kernel for one variant
tag=$1; shift
env "$@" ncu --metrics launch__occupancy_limit_shared_mem,launch__occupancy_limit_registers,launch__shared_mem_config_size,launch__registers_per_thread,sm__warps_active.avg.per_cycle_active,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts.sum,dram__bytes_read.sum,dram__bytes_write.sum \
  --clock-control none -k regex:k_advance_cc1_2d -s 6 -c 3 --csv --log-file gpurun_out/occ_$tag.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-collisions --no-c4 --no-mass-matrix > /dev/null 2> gpurun_out/occ_$tag.err
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/occ_$tag.csv')) if len(r)>10]
h=rows[0]; 
by={}
for r in rows[1:]:
    d=dict(zip(h,r)); by.setdefault(d['ID'],{})[d['Metric Name']]=d['Metric Value']; by[d['ID']]['k']=d['Kernel Name'][:60]
for i,m in by.items():
    print('$tag', i, m.pop('k'), {k.split('.')[0].replace('launch__','').replace('l1tex__data_pipe_',''):v for k,v in m.items()})
PY
