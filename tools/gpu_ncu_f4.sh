#!/bin/bash
# GPU-box script: ncu --set full capture of the one-kernel blur level (sigma / config from the environment).
export PROF_REPS=2
SIG=${NCU_SIGMA:-1.2263}
CFG=${NCU_CFG:-}
OUT=${NCU_OUT:-r2_f4}
PROF_SIGMAS=$SIG python tools/prof_levels.py "$CFG" > gpurun_out/plain_$OUT.log 2>&1 &&
PROF_SIGMAS=$SIG ncu --set full --clock-control none --import-source on --warp-sampling-interval 0 --kernel-name-base demangled -k regex:blur_ -s 6 -c 1 \
    -o gpurun_out/$OUT -f python tools/prof_levels.py "$CFG" > gpurun_out/ncu_$OUT.log 2>&1
tail -n 4 gpurun_out/plain_$OUT.log; tail -n 2 gpurun_out/ncu_$OUT.log
