#!/bin/bash
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/lp_sscd.csv python tools/layer_profile.py run sscd 256 > /dev/null 2>&1
DCR_CONV_NO_HALO=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/lp_sscd_nohalo.csv python tools/layer_profile.py run sscd 256 > /dev/null 2>&1
