#!/bin/bash
export PROF_CONTEXTS=${PROF_CONTEXTS:-6}
run() { env "$@" python tools/prof_batch.py 2>&1 | grep "^batch"; }
run S3D_F4_CTAS=148
run S3D_F4_CTAS=148 PROF_CONTEXTS=8
run S3D_F4_TY=32 S3D_F4_CTAS=148
run S3D_F4_CTAS=592
run S3D_F4_CTAS=1184
run S3D_F4_MIN_VOXELS=0
run PROF_CONTEXTS=2
run PROF_CONTEXTS=1
