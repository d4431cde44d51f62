#!/bin/bash
# where the end-to-end time of a config-5 shard goes (host time stamps of simulation(), kernel sum from the library)
mkdir -p gpurun_out
DISIMPY_B200_TRACE=1 CONFIG5_CONTAINMENT=0 timeout 600 python tools/config5.py > gpurun_out/config5_r02_ag.log 2>&1
cat gpurun_out/config5_r02_ag.log | cut -c1-1500
