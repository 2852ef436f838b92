#!/bin/bash
export PROF_CONTEXTS=${PROF_CONTEXTS:-6}
run() { env "$@" python tools/prof_batch.py 2>&1 | grep "^batch"; }
run S3D_TAIL_BLOCKS=1,2,3
run S3D_TAIL_BLOCKS=2,4,6
run S3D_TAIL_BLOCKS=6,12,20
run S3D_TAIL_BLOCKS=1,1,2
run PROF_CONTEXTS=3
run PROF_CONTEXTS=12
run S3D_DETECT_CTAS=0
run S3D_DETECT_CTAS=2
