#!/bin/bash
# GPU-box script: ncu --set full captures of the blur kernels (one-kernel level at 7 taps, x+y / z kernels at 17 taps).
# Each capture follows a plain run of the same command (profiling recipe).  Reports land in gpurun_out/.
export PROF_REPS=2
PROF_SIGMAS=1.2263 python tools/prof_levels.py "" > gpurun_out/plain1.log 2>&1 &&
PROF_SIGMAS=1.2263 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:blur_f4 -s 6 -c 1 \
    -o gpurun_out/r2_f4_r3 -f python tools/prof_levels.py "" > gpurun_out/ncu1.log 2>&1
PROF_SIGMAS=3.09 python tools/prof_levels.py "S3D_F4_MAXR=0" > gpurun_out/plain2.log 2>&1 &&
PROF_SIGMAS=3.09 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:blur_ -s 12 -c 2 \
    -o gpurun_out/r2_xy2z2_r8 -f python tools/prof_levels.py "S3D_F4_MAXR=0" > gpurun_out/ncu2.log 2>&1
tail -4 gpurun_out/plain1.log gpurun_out/plain2.log; tail -3 gpurun_out/ncu1.log gpurun_out/ncu2.log
