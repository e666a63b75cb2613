#!/bin/bash
# build_variant.sh NAME "-DFLAG=.. ..." [file.cu ...]: builds build/ab/NAME.so with the given extra flags applied to
# the listed sources (default linattn.cu); the other objects are re-used from the main build.  Use with DQ_B200_LIB.
set -e
cd "$(dirname "$0")/../diffusion-deconvolution-dia-msms-data_b200/csrc"
name=$1; flags=$2; shift 2 || true
files=${@:-linattn.cu}
mkdir -p ../../build/ab/$name
objs=""
for f in conv.cu conv_fused.cu linattn.cu linattn_tc.cu gemm_tcgen05.cu sched.cu small.cu mid.cu optim.cu data.cu; do
  if [[ " $files " == *" $f "* ]]; then
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden $flags -c $f -o ../../build/ab/$name/${f%.cu}.o
    objs="$objs ../../build/ab/$name/${f%.cu}.o"
  else
    objs="$objs ../build/${f%.cu}.o"
  fi
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../build/ab/$name.so $objs
echo built build/ab/$name.so
