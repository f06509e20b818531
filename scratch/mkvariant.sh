#!/bin/bash
# scratch/mkvariant.sh <name> <src.cu> [nvcc -D flags...]: build scratch/libs/<name>.so with one source recompiled
set -e
name=$1; src=$2; shift 2
P=cs-304-speech-recognition-code_b200
mkdir -p scratch/libs scratch/obj
base=$(basename $src .cu)
nvcc -I cs-304-speech-recognition-code_b200/csrc -I include -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v "$@" -c $src -o scratch/obj/${name}_${base}.o 2>&1 | grep -A2 -E "mel_r_kernelIf|ceps_kernel|error" | grep -v "^--" | head -40
objs=""
for o in $P/lib/*.o; do b=$(basename $o .o); if [ "$b" == "$base" ] || [ "$b" == "${base%_v2}" ]; then continue; fi; objs="$objs $o"; done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o scratch/libs/$name.so $objs scratch/obj/${name}_${base}.o -lcudart
echo built scratch/libs/$name.so
