#!/bin/bash
# Build a variant of libse_b200.so with extra nvcc flags for the tensor-core kernels (A/B and role-profile runs on one
# GPU box):   tools/build_variant.sh NAME -DSE_GEMM_PROFILE=1 -DSE_GRU_PROFILE=1 -DSE_ENC_PROFILE=1  ->  variants/libse_NAME.so
# The other objects come from the regular build (speech_enhancement_mi_b200/build/*.o).
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p variants
python -m speech_enhancement_mi_b200.build > /dev/null
for f in gemm_tc gru_wave enc_tc; do
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC "$@" \
    -c -o variants/${f}_$name.o speech_enhancement_mi_b200/csrc/$f.cu &
done
wait
objs=$(ls speech_enhancement_mi_b200/build/*.o | grep -v -E "/(gemm_tc|gru_wave|enc_tc)\.o")
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o variants/libse_$name.so $objs variants/gemm_tc_$name.o variants/gru_wave_$name.o variants/enc_tc_$name.o
rm -f variants/*_$name.o
echo variants/libse_$name.so
