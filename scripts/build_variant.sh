#!/bin/bash
# developer helper: build libmre_b200 with extra -D flags into build/variants/libmre_<name>.so (A/B timing via MRE_B200_LIB)
set -e
name=$1; shift
d=$(dirname $0)/../multimodal-relation-extrapolation_b200
mkdir -p $d/build/variants
cd $d/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC,-O2 -shared "$@" \
  -o ../build/variants/libmre_$name.so index.cpp tma_host.cpp abi.cu transe_rank.cu metrics.cu sampler.cu train_step.cu bilinear_rank.cu zsl_rank.cu peer.cu project.cu rotate_rank.cu index_build.cu
