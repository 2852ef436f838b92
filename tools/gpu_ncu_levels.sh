#!/bin/bash
# GPU-box script: ncu --set full capture of every blur level of the octave-0 schedule at MNI size (stage-level call,
# cold L2), one launch per kernel, for profiles/r2_ncu_levels.ncu-rep and the DRAM traffic bench.py quotes.
export PROF_REPS=1 NCU_ONE=1
python tools/prof_levels.py "" > gpurun_out/plain_levels.log 2>&1 &&
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:blur_ \
    -o gpurun_out/r2_ncu_levels -f python tools/prof_levels.py "" > gpurun_out/ncu_levels.log 2>&1
tail -n 4 gpurun_out/plain_levels.log; tail -n 2 gpurun_out/ncu_levels.log
