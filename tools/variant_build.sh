#!/bin/bash
# Builds a tuning variant of the engine: tools/variant_build.sh NAME [-DFLAG ...] -> build/NAME/libaddapt_b200_NAME.so
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p build/$name
ARCH="-gencode arch=compute_100a,code=sm_100a"
FLAGS="$ARCH -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-Wall,-Wno-unused-function $*"
PAR=$(realpath params/rna_turner2004_reconstructed.par)
for f in engine fold_kernels design_kernels; do
  nvcc $FLAGS -DADP_DEFAULT_PAR="\"$PAR\"" -c addapt_b200/csrc/$f.cu -o build/$name/$f.o &
done
nvcc $FLAGS -c addapt_b200/csrc/energy_model.cc -o build/$name/energy_model.o &
wait
nvcc $ARCH -shared -o build/$name/libaddapt_b200_$name.so build/$name/*.o -cudart shared -Xcompiler -static-libstdc++,-static-libgcc -Xlinker --exclude-libs,ALL
