#!/bin/bash
# usage: tools/build_variant.sh <out.so> [extra nvcc flags...]   -- builds gaz_net.cu of the working tree with extra flags into tests/_emul/<out.so>
set -e
out=$1; shift
d=$(mktemp -d)
mkdir -p $d/a/b/csrc $d/a/include
cp grok_alpha_zero_b200/csrc/*.cu grok_alpha_zero_b200/csrc/*.cuh grok_alpha_zero_b200/csrc/*.h $d/a/b/csrc/
cp include/* $d/a/include/
(cd $d/a/b/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC "$@" -c -o gaz_net.o gaz_net.cu)
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o tests/_emul/$out $d/a/b/csrc/gaz_net.o grok_alpha_zero_b200/csrc/gaz_engine.o
rm -rf $d
