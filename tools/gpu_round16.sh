#!/bin/bash
mkdir -p gpurun_out
for K in sscd vit; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/lp_$K.csv python tools/layer_profile.py run $K 256 > /dev/null 2>&1
done
