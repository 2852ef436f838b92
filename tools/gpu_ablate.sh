#!/bin/bash
# GPU-box script: batch-throughput ablation (tools/prof_batch.py under S3D_PROF_SKIP / PROF_MAX_OCT / blur knobs)
export PROF_CONTEXTS=${PROF_CONTEXTS:-6}
run() { env "$@" python tools/prof_batch.py 2>&1 | grep "^batch"; }
run X=1
run S3D_PROF_SKIP=4
run S3D_PROF_SKIP=12
run S3D_PROF_SKIP=1
run S3D_PROF_SKIP=3
run S3D_PROF_SKIP=3 PROF_MAX_OCT=1
run S3D_F4_MAXR=0
run S3D_F4_TY=32
run S3D_F4_MAXR=4
