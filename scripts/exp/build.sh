#!/bin/bash
# usage: build.sh "<RUN(...) list>"  -> pairform binary + pairform.sass
set -e
cd "$(dirname "$0")"
echo "${1:-RUN(0,256,2,2,2) RUN(1,256,2,2,2) RUN(2,256,2,2,2)}" > variants.inc
nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -lineinfo ${NVCC_EXTRA} -o pairform pairform.cu
cuobjdump -sass pairform > pairform.sass
