#!/bin/bash
# round-2 sweep: variants of the 2D CC1 tile kernel (short device-resident C3 bench each); usage: tools/sweep_r2a.sh "ENV=.. ENV=.." ...
mkdir -p gpurun_out
for cfg in "$@"; do
  tag=$(echo "$cfg" | tr ' =' '__')
  tools/quick_bench.sh $tag $cfg
done
