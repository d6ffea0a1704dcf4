#!/bin/bash
# tools/build_variant.sh NAME "-DFLAG=1 ..." : variants/NAME/libggb_cuda.so = the library with decoder.cu rebuilt under the flags.
# Knock-out flags of decoder.cu (timing experiments, wrong results): -DGGB_KO_STREAM (no weight copies), -DGGB_KO_COMPUTE
# (no term arithmetic), -DGGB_KO_FADD (one chain add per round).  tools/run_variants.sh NAME... benches them on the GPU box.
set -e
cd "$(dirname "$0")/.."
name=$1; flags=$2
mkdir -p variants/$name
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 $flags -fmad=false -prec-div=true -prec-sqrt=true -ftz=false \
  -Xcompiler -fPIC,-ffp-contract=off,-fno-fast-math -I include -I /usr/include -c gemma.ggml_b200/csrc/decoder.cu -o variants/$name/decoder.o 2>&1 | grep -v "offline compilation" || true
objs=$(ls gemma.ggml_b200/build/*.o | grep -v host_ | grep -v /decoder.o)
/usr/local/cuda/bin/nvcc -shared -o variants/$name/libggb_cuda.so $objs variants/$name/decoder.o -lcudart -ldl 2>&1 | grep -v "offline compilation" || true
ls -la variants/$name/libggb_cuda.so
