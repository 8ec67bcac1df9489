#!/bin/bash
# tools/build_k19_variant.sh NAME -DSKY_K19_...   -> skyrendering_b200/csrc/variant_NAME.so (A/B experiments; SKYB200_LIB selects it)
set -e
cd "$(dirname "$0")/../skyrendering_b200/csrc"
name=$1; shift
nvcc -O3 -std=c++20 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -ccbin /usr/bin/g++ --expt-relaxed-constexpr -use_fast_math "$@" -Xptxas -v -dc -o /tmp/pathtrace_$name.o pathtrace.cu 2> /tmp/pathtrace_$name.log
grep -A2 "k19_path_traceILi3ELb0ELi1ELb0" /tmp/pathtrace_$name.log | grep -o "Used [0-9]* registers\|[0-9]* bytes spill stores" | tr '\n' ' '; echo " <- $name"
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o variant_$name.so atmosphere.o composite.o noise.o cloud.o /tmp/pathtrace_$name.o api.o -cudart static
